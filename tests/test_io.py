"""Checkpoint directory reader (mlff_b200/io.py, mirror of mlff/io/checkpoint.py:12-56): step selection, the orbax-free
params.npz form, the loud failure without orbax, and the orbax path itself driven through a stand-in module that records
what the reference's loader would ask orbax for."""
import os
import sys
import types

import numpy as np
import pytest

from mlff_b200 import io as mio
from mlff_b200 import params as P
from mlff_b200.config import So3kratesConfig

CFG = So3kratesConfig(F=32, n_layer=2, degrees=(1, 2), n_heads=4)


def _ckpt_dirs(tmp_path):
    for name in ('ckpt_3', 'ckpt_12', 'ckpt_7', 'logs', 'ckpt_old_1'):
        os.makedirs(tmp_path / name / 'state')
    (tmp_path / 'ckpt_99').write_text('a file, not a step directory')
    return str(tmp_path)


def test_latest_step_follows_the_reference_rule(tmp_path):
    d = _ckpt_dirs(tmp_path)
    assert mio.latest_step(d) == 12
    assert mio.latest_step(str(tmp_path / 'logs')) is None


def test_params_npz_round_trip(tmp_path):
    prm = P.init_params(CFG, 3)
    mio.save_params_npz(str(tmp_path), prm)
    got = mio.load_params_from_ckpt_dir(str(tmp_path), CFG)
    assert set(got) == set(prm)
    for k in prm:
        np.testing.assert_array_equal(got[k], prm[k])


def test_orbax_checkpoint_without_orbax_fails_loudly(tmp_path, monkeypatch):
    d = _ckpt_dirs(tmp_path)
    monkeypatch.setitem(sys.modules, 'orbax', None)
    monkeypatch.setitem(sys.modules, 'orbax.checkpoint', None)
    with pytest.raises(ImportError, match='orbax'):
        mio.load_params_from_ckpt_dir(d, CFG)
    with pytest.raises(FileNotFoundError):
        mio.load_params_from_ckpt_dir(str(tmp_path / 'logs'), CFG)


def test_orbax_path_restores_valid_params_of_the_latest_step(tmp_path, monkeypatch):
    """With `orbax.checkpoint` importable the loader does what checkpoint.py:39-56 does: Checkpointer(PyTreeCheckpointHandler())
    .restore(<dir>/ckpt_<max>/state, item=None)['valid_params'] -> flat names through the flax-path map."""
    d = _ckpt_dirs(tmp_path)
    prm = P.init_params(CFG, 5)
    asked = {}

    class Handler:
        pass

    class Checkpointer:
        def __init__(self, handler):
            assert isinstance(handler, Handler)

        def restore(self, path, item=None):
            asked['path'], asked['item'] = str(path), item
            return {'valid_params': P.to_flax_tree(CFG, prm)['params'], 'params': None, 'step': 12}

    oc = types.ModuleType('orbax.checkpoint')
    oc.Checkpointer, oc.PyTreeCheckpointHandler = Checkpointer, Handler
    ob = types.ModuleType('orbax')
    ob.checkpoint = oc
    monkeypatch.setitem(sys.modules, 'orbax', ob)
    monkeypatch.setitem(sys.modules, 'orbax.checkpoint', oc)
    got = mio.load_params_from_ckpt_dir(d, CFG)
    assert asked['path'].endswith(os.path.join('ckpt_12', 'state')) and asked['item'] is None
    for k, v in prm.items():
        if k.startswith('energy/per_atom_'):
            continue                                  # constants of the Energy module: from hyperparameters.json, not the tree
        np.testing.assert_array_equal(got[k], v)


def test_calculator_from_checkpoint_directory(tmp_path, monkeypatch):
    """mlffCalculator.create_from_ckpt_dir (mlff/md/calculator.py:31-52): hyperparameters.json + scales.json + parameters;
    the energy scale multiplies the model output, add_energy_shift adds sum_atoms per_atom_shift[z] on top (not scaled,
    mlff_potential.py:80-91).  The CUDA engine is replaced by a recording stand-in: this is the host logic only."""
    import json
    import torch
    from mlff_b200 import calculator as C
    prm = P.init_params(CFG, 3)
    mio.save_params_npz(str(tmp_path), prm)
    (tmp_path / 'hyperparameters.json').write_text(json.dumps(CFG.to_hyperparameters()))
    shift = np.zeros(10)
    shift[[1, 6, 8]] = [-0.5, -37.8, -75.0]
    (tmp_path / 'scales.json').write_text(json.dumps({'energy': {'scale': [2.5], 'per_atom_shift': shift.tolist()},
                                                      'force': {'scale': [2.5]}}))
    seen = {}

    class StandIn:
        device = torch.device('cpu')

        def __init__(self, cfg, device, flags):
            seen['cfg'], seen['flags'] = cfg, flags
            self._status = torch.zeros(1, dtype=torch.int32)

        def load_params(self, flat):
            seen['params'] = flat

        def neighbor_list(self, pos, cutoff, capacity, cell=None, pbc=None, mode=None):
            ii = torch.full((capacity,), -1, dtype=torch.int32)
            return ii, ii.clone(), torch.zeros(capacity, 3, dtype=torch.int32), torch.tensor([0])

        def energy_forces(self, R, z, ii, jj, cell=None, off=None, want_stress=False):
            return torch.tensor([[1.2]]), torch.ones(1, R.shape[1], 3), None

        def raise_for_status(self, s):
            return 0

    monkeypatch.setattr(C, 'So3kEngine', StandIn)
    monkeypatch.setattr(C.mlffCalculator, '_pinned', lambda self, name, arr: torch.from_numpy(arr))
    z = np.array([6, 1, 1, 8])
    R = np.zeros((4, 3))
    calc = C.mlffCalculator.create_from_ckpt_dir(str(tmp_path), flags=3)
    assert seen['flags'] == 3 and seen['cfg'].F == CFG.F and tuple(seen['cfg'].degrees) == tuple(CFG.degrees)
    assert set(seen['params']) == set(prm) and not calc.calculate_stress
    res = calc.calculate(R, z)
    assert abs(res['energy'] - 2.5 * 1.2) < 1e-6 and np.allclose(res['forces'], 2.5)
    shifted = C.mlffCalculator.create_from_ckpt_dir(str(tmp_path), add_energy_shift=True, calculate_stress=True)
    assert shifted.calculate_stress
    assert abs(shifted.calculate(R, z)['energy'] - (2.5 * 1.2 - 37.8 - 0.5 - 0.5 - 75.0)) < 1e-5
