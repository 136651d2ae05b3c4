"""mlff_b200.inference (evaluate_model + metrics) against values recorded from the reference's own
mlff/inference/evaluation.py (tests/golden/make_reference_evaluation_golden.py), and its host-side contracts."""
import importlib.util
import os
from functools import partial

import numpy as np
import torch

from conftest import golden
from mlff_b200 import inference as inf

HERE = os.path.dirname(os.path.abspath(__file__))


def _generator():
    spec = importlib.util.spec_from_file_location('make_reference_evaluation_golden',
                                                  os.path.join(HERE, 'golden', 'make_reference_evaluation_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                    # defines obs_fn / make_data; touches /root/reference only in main()
    return mod


def test_evaluate_model_matches_reference_source():
    g = golden('reference_evaluation.npz')
    gen = _generator()
    params, inputs, targets = gen.make_data()
    metric_fn = {'mae': inf.mae_metric, 'mse': inf.mse_metric, 'rmse': inf.rmse_metric, 'r2': inf.r2_metric,
                 'mae_pad0': partial(inf.mae_metric, pad_value=0.0), 'r2_pad0': partial(inf.r2_metric, pad_value=0.0)}
    for bs in (5, 23):
        metrics, rec = inf.evaluate_model(params, gen.obs_fn, (inputs, targets), batch_size=bs, metric_fn=metric_fn,
                                          verbose=False)
        n_used = int(g[f'bs{bs}/n_used'])
        assert len(rec['predictions']['E']) == n_used == (23 // bs) * bs          # the incomplete batch is left out
        assert len(rec['inputs']['R']) == n_used and len(rec['targets']['F']) == n_used
        assert sorted(rec['predictions']) == list(g[f'bs{bs}/keys'])
        np.testing.assert_allclose(rec['predictions']['E'], g[f'bs{bs}/pred_E'], rtol=1e-13)
        for m, d in metrics.items():
            assert set(d) == {'E', 'F'}                                          # joint keys only ('extra' has no target)
            for k, v in d.items():
                np.testing.assert_allclose(v, g[f'bs{bs}/{m}/{k}'], rtol=1e-12, err_msg=f'{bs} {m} {k}')


def test_evaluate_model_accepts_device_style_outputs_and_no_targets():
    gen = _generator()
    params, inputs, targets = gen.make_data()
    tens = lambda p, x: {k: torch.as_tensor(v) for k, v in gen.obs_fn(p, x).items()}      # what a torch obs_fn returns
    metrics, rec = inf.evaluate_model(params, tens, (inputs, None), batch_size=4, metric_fn={'mae': inf.mae_metric},
                                      verbose=False)
    assert metrics == {} and rec['targets'] is None
    assert isinstance(rec['predictions']['F'], np.ndarray) and rec['predictions']['F'].shape == (20, 6, 3)
    ref = gen.obs_fn(params, {k: v[:20] for k, v in inputs.items()})
    np.testing.assert_allclose(rec['predictions']['E'], ref['E'], rtol=1e-13)


def test_metrics_on_small_numbers():
    p = np.array([1.0, 2.0, 3.0, 4.0])
    t = np.array([1.5, np.nan, 0.0, 3.0])
    assert inf.mae_metric(p, t) == (0.5 + 3.0 + 1.0) / 3
    assert inf.mae_metric(p, t, pad_value=0.0) == (0.5 + 1.0) / 2
    assert inf.mse_metric(p, t, pad_value=0.0) == (0.25 + 1.0) / 2
    assert inf.rmse_metric(p, t, pad_value=0.0) == np.sqrt((0.25 + 1.0) / 2)
    assert np.isnan(inf.mae_metric(p, t, ignore_nan=False))
    x = np.linspace(0, 1, 11)
    assert abs(inf.r2_metric(x, x) - 1.0) < 1e-15 and abs(inf.r2_metric(np.zeros(11), x)) < 1e-15
