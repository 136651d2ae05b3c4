"""Batch evaluation around the training / evaluation seam: the host-side mirror of mlff/inference/evaluation.py.

  evaluate_model(params, obs_fn, data, batch_size, metric_fn)   evaluation.py:16-109
  mae_metric / mse_metric / rmse_metric / r2_metric             evaluation.py:112-223

`obs_fn` is what `mlff_b200.nn.get_obs_and_force_fn(net)` returns (or any callable with that signature): it receives the
parameters and a dict of inputs with a leading batch axis and returns a dict of predictions (numpy arrays or torch
tensors, host or device).  Everything here is host bookkeeping -- slicing the data into full batches, concatenating the
predictions, reducing them with the metric functions; the arithmetic of the model stays behind the C ABI.
"""
from __future__ import annotations

import logging
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np

Array = Any


def _host(a) -> np.ndarray:
    if hasattr(a, 'detach'):                       # torch tensor, on any device
        return a.detach().cpu().numpy()
    return np.asarray(a)


def evaluate_model(params, obs_fn: Callable, data: Tuple[Dict, Optional[Dict]], batch_size: int,
                   metric_fn: Optional[Dict[str, Callable]] = None, verbose: bool = True) -> Tuple[Dict, Dict]:
    """Run `obs_fn` over `data = (inputs, targets or None)` in batches of `batch_size` and reduce with `metric_fn`
    ({metric name: fn(prediction=, target=)}).  As in the reference, the incomplete batch at the end of the data is LEFT OUT
    (with a warning) and inputs / targets are cut to the evaluated length; metrics are computed for the keys that
    predictions and targets share.  Returns (metrics {name: {key: value}}, {'inputs', 'predictions', 'targets'})."""
    inputs, targets = data
    n_data = len(next(iter(inputs.values())))
    n_batches = n_data // batch_size
    n_used = n_batches * batch_size
    if n_used != n_data:
        logging.warning('Number of data points modulo `batch_size` is unequal to zero. Left out %d data points at the end '
                        'of the data.', n_data - n_used)
    if verbose:
        logging.info('Model evaluation: %d data points, %d batches of %d', n_used, n_batches, batch_size)
    chunks: Dict[str, list] = {}
    for b in range(n_batches):
        sl = slice(b * batch_size, (b + 1) * batch_size)
        out = obs_fn(params, {k: v[sl] for k, v in inputs.items()})
        for k, v in out.items():
            chunks.setdefault(k, []).append(_host(v))
    predictions = {k: np.concatenate(v, axis=0) for k, v in chunks.items()}
    inputs = {k: v[:n_used] for k, v in inputs.items()}
    if targets is not None:
        targets = {k: v[:n_used] for k, v in targets.items()}
    metrics: Dict[str, Dict] = {}
    if metric_fn is not None and targets is not None:
        joint = [k for k in predictions if k in targets]
        if set(predictions) != set(targets):
            logging.warning('The model predicts %s, the data holds %s: evaluating the joint keys %s only.',
                            sorted(predictions), sorted(targets), joint)
        for name, fn in metric_fn.items():
            metrics[name] = {k: fn(prediction=predictions[k], target=_host(targets[k])) for k in joint}
    return metrics, {'inputs': inputs, 'predictions': predictions, 'targets': targets}


def _select(prediction: Array, target: Array, pad_value: Optional[float], ignore_nan: bool):
    """Flattened (prediction, target) restricted to the entries that count: target != pad_value, target not NaN."""
    p, t = _host(prediction).reshape(-1), _host(target).reshape(-1)
    keep = np.ones(len(t), dtype=bool)
    if pad_value is not None:
        keep &= t != pad_value
    if ignore_nan:
        keep &= ~np.isnan(t)
    return p[keep], t[keep]


def mae_metric(prediction: Array, target: Array, pad_value: Optional[float] = None, ignore_nan: bool = True):
    """Mean absolute error over the entries that count (evaluation.py:112-140)."""
    p, t = _select(prediction, target, pad_value, ignore_nan)
    return np.abs(p - t).mean()


def mse_metric(prediction: Array, target: Array, pad_value: Optional[float] = None, ignore_nan: bool = True):
    """Mean squared error over the entries that count (evaluation.py:143-171)."""
    p, t = _select(prediction, target, pad_value, ignore_nan)
    return ((p - t) ** 2).mean()


def rmse_metric(prediction: Array, target: Array, pad_value: Optional[float] = None, ignore_nan: bool = True):
    """Root of mse_metric; like the reference (evaluation.py:174-189) it always ignores NaN targets."""
    return np.sqrt(mse_metric(prediction=prediction, target=target, pad_value=pad_value))


def r2_metric(prediction: Array, target: Array, pad_value: Optional[float] = None, ignore_nan: bool = True) -> float:
    """R2 = 1 - (std(target - prediction) / std(target))^2 over the entries that count (evaluation.py:192-223)."""
    p, t = _select(prediction, target, pad_value, ignore_nan)
    return 1 - (np.std(t - p) / np.std(t)) ** 2
