"""Golden values from the reference's own mlff/data/preprocessing.py (get_per_atom_shift :11-51, with scikit-learn's
DictVectorizer as shipped) and mlff/data/data.py (DataTuple), executed unmodified from /root/reference.
tests/test_data_helpers.py replays the same inputs through mlff_b200.dataprep.

    python tests/golden/make_reference_data_golden.py               (build container only; needs /root/reference)
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as rs  # noqa: E402


def make_inputs():
    rng = np.random.default_rng(9)
    z = rng.choice([1, 6, 7, 8], size=(40, 9))
    z[::3, 6:] = 0                                        # differently sized structures, padded with 0
    true = np.zeros(9)
    true[[1, 6, 7, 8]] = [-0.5, -37.8, -54.6, -75.0]
    q = np.take(true, z).sum(-1) + rng.normal(scale=0.05, size=40)
    return z, q


def main():
    rs.install()
    pkg = types.ModuleType('mlff.data')
    pkg.__path__ = []
    sys.modules['mlff.data'] = pkg
    pre = rs.load('mlff.data.preprocessing')
    dat = rs.load('mlff.data.data')
    z, q = make_inputs()
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out['shifts_pad0'], out['q_scaled_pad0'] = pre.get_per_atom_shift(z, q, pad_value=0)
        zz = np.where(z == 0, 1, z)                        # no padding: every entry is an element
        out['shifts_nopad'], out['q_scaled_nopad'] = pre.get_per_atom_shift(zz, q)
    pk = {'energy': 'E', 'force': 'F', 'atomic_position': 'R', 'atomic_type': 'z'}
    dt = dat.DataTuple(inputs=['atomic_position', 'atomic_type'], targets=['energy', 'force'], prop_keys=pk)
    i, t = dt({'R': 1, 'z': 2, 'E': 3, 'F': 4, 'other': 5})
    out['datatuple_inputs'], out['datatuple_targets'] = np.array(sorted(i)), np.array(sorted(t))
    np.savez_compressed(os.path.join(HERE, 'reference_data.npz'), **out)
    print('wrote reference_data.npz; shifts', out['shifts_pad0'])


if __name__ == '__main__':
    main()
