"""Checkpoint directory -> parameters: the host-side mirror of mlff/io/checkpoint.py:12-56.

The reference stores its TrainState with orbax (`orbax-checkpoint == 0.5.23`, the one pin of its setup.py) as
``<ckpt_dir>/ckpt_<n>/state`` and reads ``state['valid_params']`` of the highest step.  orbax (and the tensorstore /
OCDBT formats underneath) is a third-party dependency that is NOT in this image, so it is imported lazily:

  * with orbax installed, ``load_params_from_ckpt_dir`` restores the reference's checkpoint through orbax's own
    ``Checkpointer(PyTreeCheckpointHandler())`` exactly as checkpoint.py:39-56 does (same step selection), and the
    flax tree goes through ``params.from_flax_tree``;
  * without it, a ``params.npz`` with the flat names of so3k.h (written by ``save_params_npz``) is read, and an orbax-only
    directory raises an ImportError that says so -- never a silent default.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Optional

import numpy as np

__STEP_PREFIX__: str = 'ckpt'


def latest_step(ckpt_dir: str) -> Optional[int]:
    """Highest n among the ``ckpt_<n>`` sub-directories (checkpoint.py:43-53), None if there is none."""
    ns = []
    for u in os.scandir(Path(ckpt_dir).resolve()):
        if u.is_dir():
            prefix_n = Path(u).stem.split('_')
            if len(prefix_n) == 2 and prefix_n[0] == __STEP_PREFIX__ and prefix_n[1].isdigit():
                ns.append(int(prefix_n[1]))
    return max(ns) if ns else None


def load_state_from_ckpt_dir(ckpt_dir: str):
    """checkpoint.py:39-56: the whole TrainState pytree of the latest step (needs orbax)."""
    step = latest_step(ckpt_dir)
    if step is None:
        raise FileNotFoundError(f'no {__STEP_PREFIX__}_<n> directory under {ckpt_dir}')
    try:
        from orbax.checkpoint import Checkpointer, PyTreeCheckpointHandler
    except ImportError as e:
        raise ImportError(f'{ckpt_dir} holds an orbax checkpoint ({__STEP_PREFIX__}_{step}/state); reading it needs the '
                          f'`orbax-checkpoint` package, which is not installed.  Convert it once where orbax is available '
                          f'(mlff_b200.io.save_params_npz) or install orbax.') from e
    ckptr = Checkpointer(PyTreeCheckpointHandler())
    return ckptr.restore(Path(ckpt_dir).resolve().absolute() / f'{__STEP_PREFIX__}_{step}/state', item=None)


def load_params_from_ckpt_dir(ckpt_dir: str, cfg) -> Dict[str, np.ndarray]:
    """Flat-name parameter dict for `cfg` from a checkpoint directory: ``params.npz`` if present, else the reference's
    orbax checkpoint (``state['valid_params']``, checkpoint.py:12-16)."""
    from .params import from_flax_tree
    npz = os.path.join(ckpt_dir, 'params.npz')
    if os.path.exists(npz):
        return dict(np.load(npz))
    state = load_state_from_ckpt_dir(ckpt_dir)
    return from_flax_tree(cfg, state['valid_params'])


def save_params_npz(ckpt_dir: str, params: Dict[str, np.ndarray]) -> str:
    """Write the flat-name parameters next to hyperparameters.json (the orbax-free form of a checkpoint)."""
    path = os.path.join(ckpt_dir, 'params.npz')
    np.savez(path, **{k: np.asarray(v, np.float32) for k, v in params.items()})
    return path
