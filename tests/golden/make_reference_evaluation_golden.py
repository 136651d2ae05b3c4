"""Golden values from the REFERENCE'S OWN evaluation driver: mlff/inference/evaluation.py (evaluate_model :16-109,
mae / mse / rmse / r2 metrics :112-223) executed unmodified from /root/reference on the numpy stand-ins of ref_shim.py
(+ a dict-only jax.tree_map), with a deterministic stand-in for the observable function.
tests/test_inference.py replays the same data through mlff_b200.inference.

    python tests/golden/make_reference_evaluation_golden.py         (build container only; needs /root/reference)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as rs  # noqa: E402


def obs_fn(params, inputs):
    """A linear 'model': E = w . sum_atoms R + b per structure, F = -w per atom (zero on padding atoms)."""
    R, z = inputs['R'], inputs['z']
    E = (R * (z != 0)[..., None]).sum(1) @ params['w'] + params['b']
    F = -np.broadcast_to(params['w'], R.shape) * (z != 0)[..., None]
    return {'E': E[:, None], 'F': F, 'extra': E[:, None] * 2}


def make_data():
    rng = np.random.default_rng(5)
    n_data, n = 23, 6
    R = rng.normal(size=(n_data, n, 3))
    z = rng.integers(1, 9, size=(n_data, n))
    z[:, -1] = 0                                          # one padding atom per structure
    params = {'w': np.array([0.3, -0.2, 0.5]), 'b': 0.1}
    E_true = obs_fn(params, {'R': R, 'z': z})['E'] + rng.normal(scale=0.1, size=(n_data, 1))
    F_true = obs_fn(params, {'R': R, 'z': z})['F'] + rng.normal(scale=0.05, size=(n_data, n, 3))
    F_true[:, -1] = 0.0                                   # padded with the pad value 0
    E_true[3, 0] = np.nan                                 # an unlabeled structure
    F_true[7, 2] = np.nan
    return params, {'R': R, 'z': z}, {'E': E_true, 'F': F_true}


def main():
    rs.install()
    sys.modules['jax'].tree_map = lambda f, *trees: {k: f(*[t[k] for t in trees]) for k in trees[0]}
    inf = types.ModuleType('mlff.inference')
    inf.__path__ = []
    sys.modules['mlff.inference'] = inf
    ev = rs.load('mlff.inference.evaluation')
    params, inputs, targets = make_data()
    out = {}
    from functools import partial
    metric_fn = {'mae': ev.mae_metric, 'mse': ev.mse_metric, 'rmse': ev.rmse_metric, 'r2': ev.r2_metric,
                 'mae_pad0': partial(ev.mae_metric, pad_value=0.0), 'r2_pad0': partial(ev.r2_metric, pad_value=0.0)}
    for bs in (5, 23):
        metrics, rec = ev.evaluate_model(params, obs_fn, (inputs, targets), batch_size=bs, metric_fn=metric_fn)
        for m, d in metrics.items():
            for k, v in d.items():
                out[f'bs{bs}/{m}/{k}'] = np.array(v)
        out[f'bs{bs}/n_used'] = np.array(len(rec['predictions']['E']))
        out[f'bs{bs}/pred_E'] = rec['predictions']['E']
        out[f'bs{bs}/keys'] = np.array(sorted(rec['predictions']))
    np.savez_compressed(os.path.join(HERE, 'reference_evaluation.npz'), **out)
    print('wrote reference_evaluation.npz:', {k: float(v) for k, v in out.items() if k.startswith('bs5/') and v.ndim == 0})


if __name__ == '__main__':
    main()
