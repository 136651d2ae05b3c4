"""Host-side checks that need no GPU: the product library builds for sm_100a, loads, and exports every symbol that
include/so3k.h declares; handle / parameter-layout / error behaviour of the non-compute entry points."""
import ctypes
import os
import re

import numpy as np
import pytest

from mlff_b200 import capi
from mlff_b200.capi import So3kLib, So3kError
from mlff_b200.config import So3kratesConfig
from mlff_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'so3k.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(so3k_[a-z_]+)\s*\(', src)))


def test_header_symbols_are_exported(cuda_lib_path):
    names = header_symbols()
    assert sorted(capi.EXPORTS) == names
    dll = ctypes.CDLL(cuda_lib_path)
    for n in names:
        assert hasattr(dll, n), n


def test_library_contains_sm100a_code(cuda_lib_path):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', cuda_lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_handle_and_param_layout(cuda_lib_path):
    lib = So3kLib(cuda_lib_path)
    assert lib.version().startswith('so3k')
    cfg = So3kratesConfig(n_layer=3)
    h = lib.create(cfg.F, cfg.n_rbf, cfg.n_layer, cfg.n_heads, list(cfg.degrees), cfg.r_cut)
    assert lib.param_count(h) == 319413          # 13200 embed + 3 * 96108 + 17689 read-out + 2 * 100 scale/shift
    off = 0
    for name, shape in P.param_shapes(cfg):      # python-side order == C-side order
        o_, c_ = lib.param_lookup(h, name)
        assert (o_, c_) == (off, int(np.prod(shape))), name
        off += c_
    ws0 = lib.workspace_bytes(h, 288, 2304)
    assert ws0 > 0
    with pytest.raises(KeyError):
        lib.param_lookup(h, 'nope')
    lib.destroy(h)
    # optional branches append their parameters per layer, in the same order on both sides, and grow the workspace
    from common import engine_flags
    cfo = So3kratesConfig(n_layer=3, layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True,
                          zbl_repulsion=True)
    h = lib.create(cfo.F, cfo.n_rbf, cfo.n_layer, cfo.n_heads, list(cfo.degrees), cfo.r_cut, flags=engine_flags(cfo))
    assert lib.param_count(h) == 319413 + 3 * (6 * 132 + 6 * 132 * 132) + 10
    off = 0
    for name, shape in P.param_shapes(cfo):
        o_, c_ = lib.param_lookup(h, name)
        assert (o_, c_) == (off, int(np.prod(shape))), name
        off += c_
    assert off == lib.param_count(h)
    assert lib.workspace_bytes(h, 288, 2304) >= ws0 + (1 + 4 * 3) * 288 * 132 * 4
    lib.destroy(h)


def test_bad_config_is_rejected(cuda_lib_path):
    lib = So3kLib(cuda_lib_path)
    with pytest.raises(So3kError):
        lib.create(130, 32, 3, 4, [1, 2, 3], 5.0)     # F % n_heads != 0  (equal_head_split)
    with pytest.raises(So3kError):
        lib.create(132, 32, 3, 4, [1, 2, 5], 5.0)     # l > 4 (spherical.py:150-152)
    with pytest.raises(So3kError):
        lib.create(132, 32, 3, 4, [1, 2, 3], -1.0)


def test_config_json_roundtrip_and_flax_tree():
    cfg = So3kratesConfig(F=36, n_layer=2, degrees=(1, 2), n_heads=4)
    cfg2 = So3kratesConfig.from_hyperparameters(cfg.to_hyperparameters())
    assert (cfg2.F, cfg2.n_layer, tuple(cfg2.degrees), cfg2.n_heads, cfg2.n_rbf, cfg2.r_cut) == (36, 2, (1, 2), 4, 32, 5.0)
    p = P.init_params(cfg, 0)
    tree = P.to_flax_tree(cfg, p)
    # flax shapes: vmapped MLP kernels carry a leading axis of size 1 (so3krates_layer.py:434-446)
    assert tree['params']['layers_0']['FeatureBlock_0']['filter_fn']['VmapMLP_0']['layers_0']['kernel'].shape == (1, 32, 36)
    back = P.from_flax_tree(cfg, tree)
    assert all(np.array_equal(back[k], p[k]) for k in p)
    # optional branches: round trip, flax names (modules are auto-numbered in call order inside So3kratesLayer)
    cfo = So3kratesConfig(F=36, n_layer=2, degrees=(1, 2), n_heads=4, layer_normalization=True, final_layer=True,
                          residual_mlp_2=True)
    cf2 = So3kratesConfig.from_hyperparameters(cfo.to_hyperparameters())
    assert (cf2.layer_normalization, cf2.final_layer, cf2.residual_mlp_1, cf2.residual_mlp_2) == (True, True, False, True)
    po = P.init_params(cfo, 0)
    to = P.to_flax_tree(cfo, po)['params']['layers_1']
    assert set(to) >= {'LayerNorm_0', 'LayerNorm_1', 'LayerNorm_2', 'ResidualMLP_0'} and 'ResidualMLP_1' not in to
    assert to['ResidualMLP_0']['Residual_0']['layers_1']['kernel'].shape == (36, 36)
    assert to['ResidualMLP_0']['Dense_0']['kernel'].shape == (36, 36) and to['LayerNorm_2']['scale'].shape == (36,)
    back = P.from_flax_tree(cfo, {'params': P.to_flax_tree(cfo, po)['params']})
    assert all(np.array_equal(back[k], po[k]) for k in po)
    # what the engine does not implement is refused: stacks whose layers differ, the non-local branches
    h = cfg.to_hyperparameters()
    h['stack_net']['layers'][0]['so3krates_layer']['layer_normalization'] = True
    with pytest.raises(NotImplementedError):
        So3kratesConfig.from_hyperparameters(h)
    h = cfg.to_hyperparameters()
    for l in h['stack_net']['layers']:
        l['so3krates_layer']['non_local_feature'] = True
    with pytest.raises(NotImplementedError):
        So3kratesConfig.from_hyperparameters(h)


def test_product_has_no_cpu_path():
    """The engine refuses to run without a CUDA device and the package never imports the oracle / emulator."""
    import torch
    pkg = os.path.join(ROOT, 'mlff_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert 'oracle' not in src.lower(), f
            # no import / load / path of the emulation layer (tests/emu, libso3k_emu, cuda_emu.h); the one permitted
            # mention is prose in capi.py's docstring ("CPU-emulated build" driven by the test-suite through the same binding)
            low = src.lower().replace('cpu-emulated build', '')
            for token in ('emu', 'so3k_oracle'):
                assert token not in low.replace('enumerate', ''), (f, token)
            import re
            assert not re.search(r'^\s*(from|import)\s+(tests|oracle|emu)', src, re.M), f
    if not torch.cuda.is_available():
        from mlff_b200.engine import So3kEngine
        with pytest.raises(RuntimeError):
            So3kEngine(So3kratesConfig(n_layer=1))


def test_hyperparameters_json_schema_matches_reference_written_files():
    """tests/golden/reference_hyperparameters.json is StackNet.__dict_repr__ (stacknet.py:89-111) of nets built by the
    reference's own init_so3krates (generated by tests/golden/make_reference_model_golden.py): the config class reads
    it and writes the identical dict back, for the default model and for one with every optional switch."""
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_hyperparameters.json')) as f:
        ref = json.load(f)
    d = So3kratesConfig.from_hyperparameters(ref['default'])
    assert (d.F, d.n_layer, tuple(d.degrees), d.n_heads, d.n_rbf, d.r_cut, d.mic) == (132, 3, (1, 2, 3), 4, 32, 5.0, False)
    assert not (d.layer_normalization or d.final_layer or d.residual_mlp_1 or d.residual_mlp_2 or d.zbl_repulsion)
    q = So3kratesConfig.from_hyperparameters(ref['optional'])
    assert (q.F, q.n_layer, tuple(q.degrees), q.n_heads, q.mic) == (64, 2, (1, 2, 3, 4), 2, True)
    assert q.layer_normalization and q.final_layer and q.residual_mlp_1 and q.residual_mlp_2
    assert q.zbl_repulsion and q.zbl_repulsion_shift == 0.37
    for cfg, tag in ((d, 'default'), (q, 'optional')):
        assert json.loads(json.dumps(cfg.to_hyperparameters())) == ref[tag], tag


def test_hyperparameters_settings_that_change_the_model_are_read_or_refused():
    """Energy's per-atom scale / shift are module constants in hyperparameters.json (observable.py:36-58), not flax
    parameters: they must reach the parameter blob; gb_* filter widths and per-layer settings outside the implemented
    path must be refused, not ignored."""
    from mlff_b200 import params as P
    cfg = So3kratesConfig(F=32, n_layer=2, degrees=(1, 2), n_heads=2)
    h = cfg.to_hyperparameters()
    scale = [0.0, 1.5, 0.0, 0.0, 0.0, 0.0, 2.5, 0.0, 0.5]
    shift = [0.0, -0.6, 0.0, 0.0, 0.0, 0.0, -38.0, 0.0, -75.0]
    h['stack_net']['observables'][0]['energy']['per_atom_scale'] = scale
    h['stack_net']['observables'][0]['energy']['per_atom_shift'] = shift
    c2 = So3kratesConfig.from_hyperparameters(h)
    assert list(c2.per_atom_scale) == scale and list(c2.per_atom_shift) == shift
    assert c2.to_hyperparameters() == h
    rng = np.random.default_rng(0)
    flat = {n: rng.normal(size=s).astype(np.float32) for n, s in P.param_shapes(c2)}
    back = P.from_flax_tree(c2, P.to_flax_tree(c2, flat))
    assert np.array_equal(back['energy/per_atom_scale'][:9], np.asarray(scale, np.float32))
    assert (back['energy/per_atom_scale'][9:] == 1).all() and (back['energy/per_atom_shift'][9:] == 0).all()
    assert np.array_equal(back['energy/per_atom_shift'][:9], np.asarray(shift, np.float32))
    for key, val in (('gb_rad_filter_features', [16, 32]), ('gb_sph_filter_features', [4, 32]), ('gb_attention', 'self_att'),
                     ('chi_cut', 0.5), ('degrees', [1, 2, 3])):
        hh = cfg.to_hyperparameters()
        hh['stack_net']['layers'][1]['so3krates_layer'][key] = val
        with pytest.raises(NotImplementedError):
            So3kratesConfig.from_hyperparameters(hh)
