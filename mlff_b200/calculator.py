"""Calculator seam: positions (+ cell) in, energy and forces out, neighbour list built on the device.

  mlffCalculator.calculate   mlff/md/calculator.py:121-147  (glp neighbour list + value_and_grad of the potential)
  CalculatorX.calculate_fn   mlff/mdx/calculator.py:42-49

Host buffers go in and come out (numpy): this is the call timed as ``e2e`` by bench.py.  Unlike the reference's glp
list (minimum-image, O(N^2)) the device cell list enumerates all periodic images (ASE semantics), so cells smaller
than 2*r_cut are handled too; overflow re-allocates with capacity_multiplier like mdx/simulate.py:96-112.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .capi import NL_ASE, STATUS_OVERFLOW
from .config import So3kratesConfig
from .engine import So3kEngine


class mlffCalculator:
    implemented_properties = ('energy', 'forces', 'stress')

    def __init__(self, cfg: So3kratesConfig, params: Dict, energy_scale: float = 1.0, capacity_multiplier: float = 1.25,
                 device: str = 'cuda:0', flags: int = 0, calculate_stress: bool = False,
                 per_atom_shift: Optional[np.ndarray] = None):
        """calculate_stress (mlff/md/calculator.py:61, 87-99): periodic calls also return 'stress' = dE/d(strain) (3,3),
        the quantity the reference differentiates through glp's strain_graph (no volume factor)."""
        from . import params as _params
        self.cfg = cfg
        self.engine = So3kEngine(cfg, device, flags)
        flat = params if ('embed/embedding' in params) else _params.from_flax_tree(cfg, params)
        self.engine.load_params(flat)
        self.energy_scale = float(energy_scale)
        self.capacity_multiplier = float(capacity_multiplier)
        self.cutoff = float(cfg.r_cut)
        self.calculate_stress = bool(calculate_stress)
        # add_energy_shift of create_from_ckpt_dir: energy += sum over atoms of per_atom_shift[z] (NOT scaled, mlff_potential.py:95-103)
        self.per_atom_shift = None if per_atom_shift is None else np.asarray(per_atom_shift, dtype=np.float64).reshape(-1)
        self.capacity: Optional[int] = None
        self.results: Dict[str, np.ndarray] = {}
        self._pin: Dict[str, torch.Tensor] = {}

    @classmethod
    def create_from_ckpt_dir(cls, ckpt_dir: str, calculate_stress: bool = False, E_to_eV: float = 1., F_to_eV_Ang: float = 1.,
                             capacity_multiplier: float = 1.25, add_energy_shift: bool = False, dtype=np.float64,
                             device: str = 'cuda:0', flags: int = 0) -> 'mlffCalculator':
        """mlff/md/calculator.py:31-52: hyperparameters.json + scales.json + the parameters of the checkpoint directory
        (mlff_b200.io).  `E_to_eV` / `F_to_eV_Ang` are accepted and, as in the reference (which only logs a reminder about
        them, :80-85), not applied; `dtype` is the reference's host dtype of the results (the device arithmetic is fp32)."""
        import json
        import os
        from .io import load_params_from_ckpt_dir
        with open(os.path.join(ckpt_dir, 'hyperparameters.json')) as f:
            cfg = So3kratesConfig.from_hyperparameters(json.load(f))
        with open(os.path.join(ckpt_dir, 'scales.json')) as f:
            scales = json.load(f)
        shift = scales['energy'].get('per_atom_shift') if add_energy_shift else None
        return cls(cfg, load_params_from_ckpt_dir(ckpt_dir, cfg),
                   energy_scale=float(np.asarray(scales['energy']['scale']).reshape(-1)[0]),
                   capacity_multiplier=capacity_multiplier, device=device, flags=flags, calculate_stress=calculate_stress,
                   per_atom_shift=shift)

    def _pinned(self, name: str, arr: np.ndarray) -> torch.Tensor:
        t = self._pin.get(name)
        if t is None or t.shape != arr.shape or t.dtype != torch.from_numpy(arr).dtype:
            t = torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype).pin_memory()
            self._pin[name] = t
        t.copy_(torch.from_numpy(arr))
        return t

    def calculate(self, positions: np.ndarray, numbers: np.ndarray, cell: Optional[np.ndarray] = None,
                  pbc=(False, False, False)) -> Dict[str, np.ndarray]:
        eng = self.engine
        dev = eng.device
        pos_h = np.ascontiguousarray(positions, dtype=np.float64)
        N = pos_h.shape[0]
        pos64 = self._pinned('pos', pos_h).to(dev, non_blocking=True)
        z = self._pinned('z', np.ascontiguousarray(numbers, dtype=np.int32)).to(dev, non_blocking=True)
        periodic = cell is not None and bool(np.any(pbc))
        cell_h = np.asarray(cell, dtype=np.float64).reshape(3, 3) if periodic else None
        if self.capacity is None:
            # first call: a generous guess from the density, refined below (allocate_fn of the reference)
            self.capacity = max(64, int(N * 96))
        while True:
            ii, jj, S, count = eng.neighbor_list(pos64, self.cutoff, self.capacity, cell=cell_h,
                                                 pbc=pbc if periodic else (False, False, False), mode=NL_ASE)
            R32 = pos64.to(torch.float32)
            cell32 = None if cell_h is None else torch.as_tensor(cell_h, dtype=torch.float32, device=dev)[None]
            want_stress = self.calculate_stress and periodic
            if want_stress:
                E, F, _, stress = eng.energy_forces(R32[None], z[None], ii[None], jj[None], cell32, S[None], want_stress=True)
            else:
                E, F, _ = eng.energy_forces(R32[None], z[None], ii[None], jj[None],
                                            cell32, None if cell32 is None else S[None])
            # one D2H for everything the host needs: energy, edge count, device status word, forces [, stress]
            parts = [E.reshape(-1).double(), count.double(), eng._status.double(), F.reshape(-1).double()]
            if want_stress:
                parts.append(stress.reshape(-1).double())
            out = torch.cat(parts).cpu()
            eng._status.zero_()
            n_edges = int(out[1].item())
            status = int(out[2].item())
            if n_edges <= self.capacity:
                eng.raise_for_status(status & ~STATUS_OVERFLOW)
                break
            self.capacity = int(np.ceil(self.capacity_multiplier * n_edges))      # overflow: re-allocate and redo
        self.n_edges = n_edges
        e_shift = 0.0 if self.per_atom_shift is None else float(self.per_atom_shift[np.asarray(numbers, dtype=np.int64)].sum())
        self.results = {'energy': float(out[0].item()) * self.energy_scale + e_shift,
                        'forces': out[3:3 + 3 * N].reshape(N, 3).numpy().astype(np.float32) * self.energy_scale}
        if want_stress:
            self.results['stress'] = out[3 + 3 * N:].reshape(3, 3).numpy().astype(np.float32) * self.energy_scale
        return self.results

    def get_potential_energy(self):
        return self.results['energy']

    def get_forces(self):
        return self.results['forces']

    def get_stress(self):
        return self.results['stress']
