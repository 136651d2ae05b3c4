"""MD-side seam: per-atom energies (and their gradient) from a Graph in the 'displacements' convention.

  MLFFPotential.create_from_ckpt_dir / potential_fn   mlff/mdx/potential/mlff_potential.py:39-117
  Graph(edges, nodes, centers, others, mask)          mlff/utils/structures.py:5
"""
from __future__ import annotations

import json
import os
from collections import namedtuple
from typing import Dict, Optional

import numpy as np
import torch

from .config import So3kratesConfig
from .engine import So3kEngine
from . import params as _params

Graph = namedtuple('Graph', ('edges', 'nodes', 'centers', 'others', 'mask'))


class MLFFPotential:
    def __init__(self, cfg: So3kratesConfig, params: Dict, energy_scale: float = 1.0,
                 per_atom_shift: Optional[np.ndarray] = None, add_shift: bool = False, device: str = 'cuda:0',
                 flags: int = 0):
        self.cfg = cfg
        self.cutoff = float(cfg.r_cut)                       # get_model_cutoff, mlff_potential.py:18-23
        self.effective_cutoff = cfg.n_layer * self.cutoff    # mlff_potential.py:71-74
        self.energy_scale = float(energy_scale)
        self.add_shift = bool(add_shift)
        self.per_atom_shift = None if per_atom_shift is None else np.asarray(per_atom_shift, np.float32)
        self.dtype = torch.float32
        self.engine = So3kEngine(cfg, device, flags)
        flat = params if ('embed/embedding' in params) else _params.from_flax_tree(cfg, params)
        self.engine.load_params(flat)
        self._shift_dev = None
        if self.add_shift and self.per_atom_shift is not None:
            self._shift_dev = torch.as_tensor(self.per_atom_shift, device=self.engine.device)

    @classmethod
    def create_from_ckpt_dir(cls, ckpt_dir: str, add_shift: bool = False, dtype=None, device: str = 'cuda:0'):
        """hyperparameters.json + scales.json as written by the reference (mlff_train_so3krates.py:472-477); parameters
        through mlff_b200.io.load_params_from_ckpt_dir: ``params.npz`` (flat names) or, where orbax is installed, the
        reference's own ``ckpt_<n>/state`` (mlff/io/checkpoint.py:12-56; mdx/potential/mlff_potential.py:58-63)."""
        from .io import load_params_from_ckpt_dir
        with open(os.path.join(ckpt_dir, 'hyperparameters.json')) as f:
            cfg = So3kratesConfig.from_hyperparameters(json.load(f))
        with open(os.path.join(ckpt_dir, 'scales.json')) as f:
            scales = json.load(f)
        params = load_params_from_ckpt_dir(ckpt_dir, cfg)
        return cls(cfg, params, energy_scale=float(np.asarray(scales['energy']['scale']).reshape(-1)[0]),
                   per_atom_shift=scales['energy'].get('per_atom_shift'), add_shift=add_shift, device=device)

    def __call__(self, graph: Graph) -> torch.Tensor:
        return self.potential_fn(graph)

    def potential_fn(self, graph: Graph) -> torch.Tensor:
        """Per-atom energies (N,), scaled (and optionally shifted) as mlff_potential.py:80-109."""
        ea, _, _ = self.engine.potential(_t(graph.edges), _t(graph.nodes), _t(graph.centers), _t(graph.others),
                                         None if graph.mask is None else _t(graph.mask), self.energy_scale,
                                         want_dE_dedges=False, want_forces=False)
        return self._shift(ea, graph)

    def value_and_grad(self, graph: Graph):
        """(sum of per-atom energies, d/d edges, forces on atoms): what the reference's calculators obtain with
        jax.value_and_grad around potential(graph).sum() (mdx/calculator.py:46-49, md/calculator.py:105-109)."""
        ea, dE, F = self.engine.potential(_t(graph.edges), _t(graph.nodes), _t(graph.centers), _t(graph.others),
                                          None if graph.mask is None else _t(graph.mask), self.energy_scale)
        ea = self._shift(ea, graph)
        return ea.double().sum(), dE, F

    def _shift(self, ea, graph):
        if self._shift_dev is not None:
            z = _t(graph.nodes).to(self.engine.device, torch.long)
            ea = ea + self._shift_dev[z]
        return ea


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
