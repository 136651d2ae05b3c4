"""Parameter handling: random initialisation, flat <-> flax-tree naming, blob packing.

The engine consumes ONE float32 blob whose order is documented in include/so3k.h.  The reference stores a flax
variable dict (``valid_params``, mlff/io/checkpoint.py:12-16); the tree paths below are the ones flax generates for
the reference's modules (derived from the module code -- a real checkpoint was not available --
and confirmed by running the reference's own model code on a tree built by `to_flax_tree`, with every leaf looked up
by flax's naming rules: tests/golden/make_reference_model_golden.py):

  params/feature_embeddings_0/Embed_0/embedding                                             (100, F)
  params/layers_t/FeatureBlock_0/filter_fn/VmapMLP_{0=rad,1=sph}/layers_{0,1}/{kernel (1,in,out), bias (1,out)}
  params/layers_t/FeatureBlock_0/attention_fn/VmapConvAttentionCoefficients_0/Dense_{0=q,1=k}/kernel (H,d,d)
  params/layers_t/FeatureBlock_0/attention_fn/VmapAttentionAggregation_0/Dense_0/kernel      (H,d,d)
  params/layers_t/GeometricBlock_0/...            (same, nl heads, no aggregation Dense)
  params/layers_t/InteractionBlock_0/MLP_0/layers_0/{kernel, bias}
  params/layers_t/LayerNorm_{0,1[,2]}/{scale, bias}                                        (layer_normalization)
  params/layers_t/ResidualMLP_k/{Residual_0/layers_{0,1}/kernel, Dense_0/kernel}           (residual_mlp_1 / _2)
  params/observables_0/MLP_0/layers_{0,1}/{kernel, bias}
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

from .config import So3kratesConfig


def param_shapes(cfg: So3kratesConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in blob order -- must agree with make_layout() in csrc/so3k_api.cu."""
    F, K, nl, H = cfg.F, cfg.n_rbf, len(cfg.degrees), cfg.n_heads
    Fs, dh, dg = int(F / 4), F // H, F // nl
    out: List[Tuple[str, Tuple[int, ...]]] = [('embed/embedding', (cfg.num_embeddings, F))]
    for t in range(cfg.n_layer):
        for blk, heads, d in (('fb', H, dh), ('gb', nl, dg)):
            pre = f'layers_{t}/{blk}/'
            out += [(pre + 'rad/layers_0/kernel', (K, F)), (pre + 'rad/layers_0/bias', (F,)),
                    (pre + 'rad/layers_1/kernel', (F, F)), (pre + 'rad/layers_1/bias', (F,)),
                    (pre + 'sph/layers_0/kernel', (nl, Fs)), (pre + 'sph/layers_0/bias', (Fs,)),
                    (pre + 'sph/layers_1/kernel', (Fs, F)), (pre + 'sph/layers_1/bias', (F,)),
                    (pre + 'q/kernel', (heads, d, d)), (pre + 'k/kernel', (heads, d, d))]
            if blk == 'fb':
                out += [(pre + 'v/kernel', (heads, d, d))]
        out += [(f'layers_{t}/int/layers_0/kernel', (F + nl, F + nl)), (f'layers_{t}/int/layers_0/bias', (F + nl,))]
        if cfg.layer_normalization:      # flax nn.LayerNorm: scale, bias (so3krates_layer.py:103-106, 153-156, 170-174)
            for k in range(3 if cfg.final_layer else 2):
                out += [(f'layers_{t}/ln_{k + 1}/scale', (F,)), (f'layers_{t}/ln_{k + 1}/bias', (F,))]
        for r, on in ((1, cfg.residual_mlp_1), (2, cfg.residual_mlp_2)):      # ResidualMLP, mlp.py:30-72 (bias-free)
            if on:
                out += [(f'layers_{t}/res_{r}/{k}/kernel', (F, F)) for k in ('layers_0', 'layers_1', 'out')]
    out += [('energy/layers_0/kernel', (F, F)), ('energy/layers_0/bias', (F,)),
            ('energy/layers_1/kernel', (F, 1)), ('energy/layers_1/bias', (1,)),
            ('energy/per_atom_scale', (cfg.num_embeddings,)), ('energy/per_atom_shift', (cfg.num_embeddings,))]
    if cfg.zbl_repulsion:            # raw ZBLRepulsion scalars (softplus applied at use), observable.py:133-142
        out += [('zbl/' + k, (1,)) for k in ZBL_NAMES]
    return out


ZBL_NAMES = ('a1', 'a2', 'a3', 'a4', 'c1', 'c2', 'c3', 'c4', 'p', 'd')
_ZBL_INIT = {'a1': 3.2, 'a2': 0.9423, 'a3': 0.4028, 'a4': 0.2016, 'c1': 0.1818, 'c2': 0.5099, 'c3': 0.2802, 'c4': 0.02817,
             'p': 0.23, 'd': 1.0 / (0.8854 * 0.5291772105638411)}


def init_params(cfg: So3kratesConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Counterpart of ``net.init(jax.random.PRNGKey(seed), inputs)``: flax defaults -- lecun_normal kernels,
    zero biases, Embed = variance_scaling(1, 'fan_in', 'normal', out_axis=0).  (The PRNG stream differs from
    JAX's; only the distributions match.)"""
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg):
        if name.endswith('/bias'):
            p[name] = np.zeros(shape, np.float32)
        elif name.endswith('/scale'):
            p[name] = np.ones(shape, np.float32)
        elif name == 'energy/per_atom_scale':
            p[name] = np.ones(shape, np.float32)
        elif name == 'energy/per_atom_shift':
            p[name] = np.zeros(shape, np.float32)
        elif name.startswith('zbl/'):
            v = _ZBL_INIT[name[4:]]
            p[name] = np.array([v + math.log(-math.expm1(-v))], np.float32)     # softplus_inverse, as the reference
        elif name == 'embed/embedding':
            p[name] = rng.normal(0.0, 1.0 / math.sqrt(shape[1]), shape).astype(np.float32)
        else:
            fan_in = shape[-2]
            # truncated normal of flax' lecun_normal is approximated by a plain normal
            p[name] = rng.normal(0.0, 1.0 / math.sqrt(fan_in), shape).astype(np.float32)
    return p


def pack_blob(cfg: So3kratesConfig, params: Dict[str, np.ndarray]) -> np.ndarray:
    chunks = []
    for name, shape in param_shapes(cfg):
        if name not in params:
            raise KeyError(f'missing parameter {name}')
        a = np.asarray(params[name], dtype=np.float32)
        if a.size != int(np.prod(shape)):
            raise ValueError(f'parameter {name}: expected shape {shape}, got {a.shape}')
        chunks.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(chunks))


# ---- flax tree <-> flat names ---------------------------------------------------------------------
def _flax_path(name: str, res_first: bool = True) -> Tuple[List[str], bool]:
    """flat name -> (path inside variables['params'], has_leading_vmap_axis).  ``res_first``: the layer has a
    ResidualMLP after the first skip connection (it is then ResidualMLP_0 and the second one ResidualMLP_1)."""
    parts = name.split('/')
    if parts[0] == 'embed':
        return ['feature_embeddings_0', 'Embed_0', 'embedding'], False
    if parts[0] == 'zbl':
        return ['observables_0', 'ZBLRepulsion_0', parts[1]], False
    if parts[0] == 'energy':
        if parts[1] in ('per_atom_scale', 'per_atom_shift'):
            return [], False
        return ['observables_0', 'MLP_0', parts[1], parts[2]], False
    lt = parts[0]
    if parts[1].startswith('ln_'):       # auto-named in call order inside So3kratesLayer.__call__
        return [lt, f'LayerNorm_{int(parts[1][3:]) - 1}', parts[2]], False
    if parts[1].startswith('res_'):      # ResidualMLP_k counts only the ResidualMLPs that exist in the layer
        k = int(parts[1][4:]) - 1 if res_first else 0
        if parts[2] == 'out':
            return [lt, f'ResidualMLP_{k}', 'Dense_0', 'kernel'], False
        return [lt, f'ResidualMLP_{k}', 'Residual_0', parts[2], 'kernel'], False
    if parts[1] == 'int':
        return [lt, 'InteractionBlock_0', 'MLP_0', parts[2], parts[3]], False
    blk = 'FeatureBlock_0' if parts[1] == 'fb' else 'GeometricBlock_0'
    if parts[2] in ('rad', 'sph'):
        return [lt, blk, 'filter_fn', 'VmapMLP_0' if parts[2] == 'rad' else 'VmapMLP_1', parts[3], parts[4]], True
    if parts[2] in ('q', 'k'):
        return [lt, blk, 'attention_fn', 'VmapConvAttentionCoefficients_0', 'Dense_0' if parts[2] == 'q' else 'Dense_1',
                'kernel'], False
    if parts[2] == 'v':
        return [lt, blk, 'attention_fn', 'VmapAttentionAggregation_0', 'Dense_0', 'kernel'], False
    raise KeyError(name)


def from_flax_tree(cfg: So3kratesConfig, variables: Dict, per_atom_scale=None, per_atom_shift=None
                   ) -> Dict[str, np.ndarray]:
    """Nested flax variable dict ({'params': ...} or the inner dict) -> flat names."""
    tree = variables['params'] if 'params' in variables else variables
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg):
        path, vm = _flax_path(name, cfg.residual_mlp_1)
        if not path:
            continue
        node = tree
        for k in path:
            node = node[k]
        a = np.asarray(node, dtype=np.float32)
        if vm:
            a = a.reshape(a.shape[1:]) if a.shape[0] == 1 and a.ndim == len(shape) + 1 else a
        out[name] = a.reshape(shape)
    # Energy's per-atom scale / shift are constants of the module, not flax parameters (observable.py:49-58): they come
    # from hyperparameters.json through the config unless the caller overrides them
    n = cfg.num_embeddings
    for key, arg, fill in (('per_atom_scale', per_atom_scale, 1.0), ('per_atom_shift', per_atom_shift, 0.0)):
        v = arg if arg is not None else getattr(cfg, key, None)
        a = np.full(n, fill, np.float32)
        if v is not None:
            v = np.asarray(v, np.float32).reshape(-1)
            if v.size > n:
                raise ValueError(f'{key} has {v.size} entries for num_embeddings = {n}')
            a[:v.size] = v                       # jnp.take(per_atom_scale, z): indexed by atomic number
        out[f'energy/{key}'] = a
    return out


def to_flax_tree(cfg: So3kratesConfig, params: Dict[str, np.ndarray]) -> Dict:
    tree: Dict = {}
    for name, shape in param_shapes(cfg):
        path, vm = _flax_path(name, cfg.residual_mlp_1)
        if not path:
            continue
        node = tree
        for k in path[:-1]:
            node = node.setdefault(k, {})
        a = np.asarray(params[name], np.float32).reshape(shape)
        node[path[-1]] = a[None] if vm else a
    return {'params': tree}
