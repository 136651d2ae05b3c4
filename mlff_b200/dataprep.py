"""The two small data helpers that sit directly in front of the training / evaluation seam.

  DataTuple            mlff/data/data.py:7-21           dataset dict -> (inputs, targets) by property keys
  get_per_atom_shift   mlff/data/preprocessing.py:11-51 least-squares per-element energy shifts (what becomes
                                                        `per_atom_shift` of the Energy read-out and `scales.json`)

Host-side numpy; the data-set class with its splitting / scaling bookkeeping (mlff/data/dataset.py) is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

Array = Any


@dataclass
class DataTuple:
    """`DataTuple(inputs, targets, prop_keys)(ds)` -> ({input key: array}, {target key: array}); the names in `inputs` /
    `targets` are property names (e.g. 'atomic_position', 'energy'), `prop_keys` maps them to the keys of the data set."""
    inputs: Sequence[str]
    targets: Sequence[str]
    prop_keys: Dict[str, str]

    def __post_init__(self):
        self.input_keys = [self.prop_keys[i] for i in self.inputs]
        self.target_keys = [self.prop_keys[t] for t in self.targets]

    def __call__(self, ds: Dict[str, Array]) -> Tuple[Dict[str, Array], Dict[str, Array]]:
        pick = lambda keys: {k: v for k, v in ds.items() if k in keys}
        return pick(self.input_keys), pick(self.target_keys)


def get_per_atom_shift(z: Array, q: Array, pad_value: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Per-element shifts x solving count(z) @ x = q in the least-squares sense: z (n_data, n) atomic numbers (padded
    with `pad_value` for differently sized structures), q (n_data) the quantity to fit.  Returns (shifts indexed BY ATOMIC
    NUMBER, length max(z) + 1, zero for absent elements and for the pad value; q - sum of its structure's shifts).
    As in the reference the correction `np.take(shifts, z)` also runs over padded entries, whose shift is zero."""
    z = np.asarray(z)
    q = np.asarray(q)
    elements = np.unique(z)
    fit = elements if pad_value is None else elements[elements != pad_value]
    counts = (z[:, :, None] == fit[None, None, :]).sum(axis=1).astype(np.float64)       # (n_data, n_elements)
    sol = np.linalg.lstsq(counts, q, rcond=None)[0]
    shifts = np.zeros(int(elements.max()) + 1)
    shifts[fit] = sol
    return shifts, q - np.take(shifts, z).sum(axis=-1)
