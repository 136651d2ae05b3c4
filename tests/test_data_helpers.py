"""mlff_b200.dataprep (DataTuple, get_per_atom_shift) against values recorded from the reference's own source
(tests/golden/make_reference_data_golden.py)."""
import importlib.util
import os

import numpy as np

from conftest import golden
from mlff_b200 import dataprep as D

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    spec = importlib.util.spec_from_file_location('make_reference_data_golden',
                                                  os.path.join(HERE, 'golden', 'make_reference_data_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.make_inputs()


def test_per_atom_shift_matches_reference_source():
    g = golden('reference_data.npz')
    z, q = _inputs()
    shifts, q_scaled = D.get_per_atom_shift(z, q, pad_value=0)
    np.testing.assert_allclose(shifts, g['shifts_pad0'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(q_scaled, g['q_scaled_pad0'], rtol=1e-8, atol=1e-9)
    assert shifts.shape == (9,) and shifts[0] == 0 and not shifts[[2, 3, 4, 5]].any()
    assert abs(shifts[6] + 37.8) < 0.1 and np.abs(q_scaled).max() < 0.3           # the fit recovers the planted shifts
    zz = np.where(z == 0, 1, z)
    shifts, q_scaled = D.get_per_atom_shift(zz, q)
    np.testing.assert_allclose(shifts, g['shifts_nopad'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(q_scaled, g['q_scaled_nopad'], rtol=1e-8, atol=1e-9)


def test_data_tuple_splits_by_property_keys():
    g = golden('reference_data.npz')
    pk = {'energy': 'E', 'force': 'F', 'atomic_position': 'R', 'atomic_type': 'z'}
    dt = D.DataTuple(inputs=['atomic_position', 'atomic_type'], targets=['energy', 'force'], prop_keys=pk)
    i, t = dt({'R': 1, 'z': 2, 'E': 3, 'F': 4, 'other': 5})
    assert sorted(i) == list(g['datatuple_inputs']) and sorted(t) == list(g['datatuple_targets'])
    assert i == {'R': 1, 'z': 2} and t == {'E': 3, 'F': 4}
    assert dt.input_keys == ['R', 'z'] and dt.target_keys == ['E', 'F']
