"""The two small data helpers that sit directly in front of the training / evaluation seam.

  DataTuple            mlff/data/data.py:7-21           dataset dict -> (inputs, targets) by property keys
  get_per_atom_shift   mlff/data/preprocessing.py:11-51 least-squares per-element energy shifts (what becomes
                                                        `per_atom_shift` of the Energy read-out and `scales.json`)

Host-side numpy; the data-set class with its splitting / scaling bookkeeping (mlff/data/dataset.py) is out of scope.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

Array = Any


class DataTuple:
    """`DataTuple(inputs, targets, prop_keys)(ds)` -> ({input key: array}, {target key: array}); the names in `inputs` /
    `targets` are property names (e.g. 'atomic_position', 'energy'), `prop_keys` maps them to the keys of the data set.
    Keys of `ds` that belong to neither side are dropped; a requested key that `ds` lacks is simply absent from the result
    (the reference filters the same way)."""

    def __init__(self, inputs: Sequence[str], targets: Sequence[str], prop_keys: Dict[str, str]):
        self.inputs, self.targets, self.prop_keys = list(inputs), list(targets), dict(prop_keys)
        self.input_keys = [prop_keys[name] for name in self.inputs]
        self.target_keys = [prop_keys[name] for name in self.targets]

    def _subset(self, ds: Dict[str, Array], wanted) -> Dict[str, Array]:
        wanted = set(wanted)
        return {key: ds[key] for key in ds if key in wanted}

    def __call__(self, ds: Dict[str, Array]) -> Tuple[Dict[str, Array], Dict[str, Array]]:
        return self._subset(ds, self.input_keys), self._subset(ds, self.target_keys)

    def __repr__(self):
        return f'DataTuple(inputs={self.inputs}, targets={self.targets})'


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
