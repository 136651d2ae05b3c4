"""Python-side mirror of the reference's model interface for the hot path, backed by libso3k.

  So3krates(...)              mlff/nn/representation/so3krates.py:13-49   (constructor arguments)
  StackNet.apply / to_json    mlff/nn/stacknet/stacknet.py:40-111
  init_stack_net(h)           mlff/nn/stacknet/stacknet.py:135-145
  get_obs_and_force_fn(net)   mlff/nn/stacknet/observable_function.py:127-149
  get_observable_fn(net)      mlff/nn/stacknet/observable_function.py:19-38

The functions returned here take ``(params, inputs)`` like the reference's jitted/vmapped ``obs_fn``; ``inputs`` uses
the reference's keys (prop_keys) and padded-batch conventions: R (B,n,3), z (B,n) with 0 = padding, idx_i / idx_j
(B,P) with -1 = padding, optional unit_cell (B,3,3) row-wise + cell_offset (B,P,3).  A single un-batched structure
(R (n,3)) is accepted too -- what the reference's un-vmapped obs_fn takes.  Arrays may be numpy (copied to the GPU) or
torch CUDA tensors (used in place); outputs follow the input kind.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from .config import So3kratesConfig, md17_property_keys
from .engine import So3kEngine
from . import params as _params


class StackNet:
    """The hot-path subset of the reference's StackNet: one GeometryEmbed, one AtomTypeEmbed, L So3kratesLayer,
    one Energy observable."""

    def __init__(self, cfg: So3kratesConfig, device: str = 'cuda:0', flags: int = 0):
        self.cfg = cfg
        self.prop_keys = dict(cfg.prop_keys)
        self.device = device
        self.flags = flags
        self._engine: Optional[So3kEngine] = None
        self._loaded_id = None
        self.input_convention = 'positions'
        self.output_convention = 'per_structure'

    # -- reference API -------------------------------------------------------------------------
    def init(self, rng_seed=0, inputs=None) -> Dict:
        """``net.init(PRNGKey, inputs)``: returns {'params': <flax-like tree>}."""
        seed = int(rng_seed) if np.isscalar(rng_seed) else int(np.asarray(rng_seed).reshape(-1)[-1])
        return _params.to_flax_tree(self.cfg, _params.init_params(self.cfg, seed))

    def apply(self, params, inputs) -> Dict:
        return get_observable_fn(self)(params, inputs)

    def __dict_repr__(self) -> Dict:
        return self.cfg.to_hyperparameters()

    def to_json(self, ckpt_dir, name='hyperparameters.json'):
        with open(os.path.join(ckpt_dir, name), 'w', encoding='utf-8') as f:
            json.dump(self.__dict_repr__(), f, ensure_ascii=False, indent=4)

    @classmethod
    def create_from_ckpt_dir(cls, ckpt_dir: str, **kw):
        with open(os.path.join(ckpt_dir, 'hyperparameters.json')) as f:
            return init_stack_net(json.load(f), **kw)

    def reset_prop_keys(self, prop_keys, sub_modules=True):
        self.prop_keys.update(prop_keys)

    def reset_input_convention(self, input_convention):
        if input_convention not in ('positions', 'displacements'):
            raise ValueError(f'{input_convention} is not a valid argument for `input_convention`.')
        self.input_convention = input_convention

    def reset_output_convention(self, output_convention):
        if output_convention not in ('per_structure', 'per_atom'):
            raise ValueError(f'{output_convention} is invalid argument for attribute `output_convention`.')
        self.output_convention = output_convention

    # -- engine plumbing ---------------------------------------------------------------------------
    def engine(self, params) -> So3kEngine:
        if self._engine is None:
            self._engine = So3kEngine(self.cfg, self.device, self.flags)
        if self._loaded_id != id(params):
            flat = params if ('embed/embedding' in params) else _params.from_flax_tree(self.cfg, params)
            self._engine.load_params(flat)
            self._loaded_id = id(params)
        return self._engine


def So3krates(prop_keys: Dict[str, str] = None, F: int = 132, n_layer: int = 6,
              so3krates_layer_kwargs: Dict = None, geometry_embed_kwargs: Dict = None, device: str = 'cuda:0',
              flags: int = 0) -> StackNet:
    """mlff/nn/representation/so3krates.py:13-49 (default embeddings / observable only)."""
    layer = {'degrees': [1, 2, 3], 'n_heads': 4}
    layer.update(so3krates_layer_kwargs or {})
    geo = {'degrees': [1, 2, 3], 'radial_basis_function': 'phys', 'n_rbf': 32, 'r_cut': 5,
           'radial_cutoff_fn': 'cosine_cutoff_fn', 'sphc': True, 'mic': False}
    geo.update(geometry_embed_kwargs or {})
    if list(layer['degrees']) != list(geo['degrees']):
        raise ValueError('layer and geometry-embedding degrees must agree')
    for k, ok in (('radial_basis_function', 'phys'), ('radial_cutoff_fn', 'cosine_cutoff_fn'), ('sphc', True)):
        if geo[k] != ok:
            raise NotImplementedError(f'{k}={geo[k]!r} is outside the hot path implemented by mlff_b200')
    cfg = So3kratesConfig(F=F, n_layer=n_layer, degrees=tuple(geo['degrees']), n_heads=int(layer['n_heads']),
                          n_rbf=int(geo['n_rbf']), r_cut=float(geo['r_cut']), mic=bool(geo['mic']),
                          prop_keys=dict(prop_keys or md17_property_keys))
    return StackNet(cfg, device=device, flags=flags)


def init_stack_net(h: Dict, device: str = 'cuda:0', flags: int = 0) -> StackNet:
    return StackNet(So3kratesConfig.from_hyperparameters(h), device=device, flags=flags)


def _to_dev(a, dev, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dev, dtype)
    return torch.as_tensor(np.asarray(a)).to(dev, dtype)


def _run(net: StackNet, params, inputs, want_forces: bool, want_stress: bool = False) -> Dict:
    pk = net.prop_keys
    eng = net.engine(params)
    dev = eng.device
    R_in = inputs[pk['atomic_position']]
    as_numpy = not isinstance(R_in, torch.Tensor)
    R = _to_dev(R_in, dev, torch.float32)
    single = R.dim() == 2
    z = _to_dev(inputs[pk['atomic_type']], dev, torch.int32)
    ii = _to_dev(inputs[pk['idx_i']], dev, torch.int32)
    jj = _to_dev(inputs[pk['idx_j']], dev, torch.int32)
    cell = off = None
    if net.cfg.mic or (pk['unit_cell'] in inputs and pk['cell_offset'] in inputs and inputs[pk['unit_cell']] is not None
                       and inputs[pk['cell_offset']] is not None):
        cell = _to_dev(inputs[pk['unit_cell']], dev, torch.float32)
        off = _to_dev(inputs[pk['cell_offset']], dev, torch.int32)
    if single:
        R, z, ii, jj = R[None], z[None], ii[None], jj[None]
        if cell is not None:
            cell, off = cell[None], off[None]
    stress = None
    if want_stress:
        if cell is None:
            raise ValueError('the stress needs `unit_cell` and `cell_offset` in the inputs')
        E, F, ea, stress = eng.energy_forces(R, z, ii, jj, cell, off, want_e_atom=(net.output_convention == 'per_atom'),
                                             want_stress=True)
    else:
        E, F, ea = eng.energy_forces(R, z, ii, jj, cell, off, want_e_atom=(net.output_convention == 'per_atom'))
    out = {}
    if net.output_convention == 'per_atom':
        out[pk.get('atomic_energy', pk['energy'])] = ea[..., None]
    else:
        out[pk['energy']] = E
    if want_forces:
        out[pk['force']] = F
    if stress is not None:
        out[pk['stress']] = stress
    if single:
        out = {k: v[0] for k, v in out.items()}
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return out


def get_observable_fn(net: StackNet, observable_key: str = None) -> Callable:
    def observable_fn(p, x):
        out = _run(net, p, x, want_forces=False)
        return out if observable_key is None else {observable_key: out[observable_key]}
    return observable_fn


def get_obs_and_force_fn(net: StackNet) -> Callable:
    """E and F = -dE/dR in one call (the reference evaluates jacrev of the energy, observable_function.py:59-61;
    here the force is the hand-written analytic backward inside libso3k)."""
    def obs_and_force_fn(p, x):
        return _run(net, p, x, want_forces=True)
    return obs_and_force_fn


def get_energy_force_stress_fn(net: StackNet) -> Callable:
    """E, F and the stress = dE/d(strain) in one call -- mlff/nn/stacknet/observable_function.py:152-186, where positions
    and cell are strained by the symmetrised eps and `jax.value_and_grad` differentiates with respect to both; here
    `so3k_stress` contracts the edge cotangents of the same analytic backward (no volume factor, as in the reference).
    The inputs need `unit_cell` and `cell_offset`; batched inputs give one (3,3) tensor per structure."""
    def energy_force_stress_fn(p, x):
        return _run(net, p, x, want_forces=True, want_stress=True)
    return energy_force_stress_fn


def get_grad_observable_fn(net: StackNet, derivatives) -> Callable:
    """Gradients of observables with respect to inputs, requested as in the reference (observable_function.py:41-95):
    each entry is `(name, (observable key, input key, op))` and the result holds `name: op(d observable / d input)`, the
    raw gradient of the (1,)-shaped energy carrying the reference's extra leading axis -- (1, n, 3) per structure -- so
    that the usual `('F', ('E', 'R', lambda y: -y.squeeze(-3)))` works unchanged.  The hot path provides exactly one
    derivative, dE/dR (the analytic force backward); anything else raises NotImplementedError."""
    pk = net.prop_keys
    for _, (obs, inp, _) in derivatives:
        if obs != pk['energy'] or inp != pk['atomic_position']:
            raise NotImplementedError(f'd {obs} / d {inp}: only d {pk["energy"]} / d {pk["atomic_position"]} is on the hot path')

    def all_grad_obs_fn(p, x):
        F = _run(net, p, x, want_forces=True)[pk['force']]
        dE_dR = -F
        lead = dE_dR[None] if dE_dR.ndim == 2 else dE_dR[:, None]         # (1, n, 3), or (B, 1, n, 3) for a batch
        return {name: op(lead) for name, (_, _, op) in derivatives}
    return all_grad_obs_fn


def get_obs_and_grad_obs_fn(net: StackNet, derivatives=None) -> Callable:
    """Observables and the requested gradients in one dict (observable_function.py:98-124)."""
    obs_fn = get_observable_fn(net)
    if derivatives is None:
        return obs_fn
    grad_fn = get_grad_observable_fn(net, derivatives)

    def obs_and_grad_fn(p, x):
        out = obs_fn(p, x)
        out.update(grad_fn(p, x))
        return out
    return obs_and_grad_fn
