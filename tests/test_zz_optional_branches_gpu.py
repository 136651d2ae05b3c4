"""Optional So3kratesLayer branches on the device: layer_normalization / final_layer / residual_mlp_1 / residual_mlp_2
(mlff/nn/layer/so3krates_layer.py:103-174, mlff/nn/mlp/mlp.py:30-72) against the CPU oracle, in the fp32 and the
tensor-core modes.  The same comparison runs on the CPU-emulated kernels in tests/test_emu_parity.py.
(This file sorts last so that `pytest -x` reaches it after the default model's suite.)"""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from common import jittered_lattice, water_like_numbers

pytestmark = pytest.mark.gpu

E_RTOL = 1e-5
F_ATOL = 1e-4


def tolerances(flags, f_max):
    """north_star's tolerances as they stand, in every arithmetic mode: energies 1e-5 relative to |E|, forces 1e-4 eV/A
    max-abs.  (The split-bf16 tensor mode carries a RELATIVE force error of 1e-5 .. 2.5e-5 -- profiles/
    r01k_optional_branch_errors.txt -- so the absolute bound presumes forces of trained-model size, a few eV/A; the
    LayerNorm test models are scaled to that with `readout_gain`, see GAIN below.)"""
    return E_RTOL, F_ATOL


# LayerNorm re-normalises the features: with unit-scale random read-out weights those models have 5-11 eV/A forces.
# The read-out gain brings them to the ~1 eV/A of the default model and of trained force fields (synthetic
# parameters with O(1) forces, so that the absolute force tolerance is meaningful).
GAIN = {'ln': 0.1, 'ln+post': 0.2, 'all': 0.2}
OPTS = {'ln': dict(layer_normalization=True),
        'ln+post': dict(layer_normalization=True, final_layer=True),
        'res1': dict(residual_mlp_1=True),
        'res2': dict(residual_mlp_2=True),
        'all': dict(layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True)}


def make(ocfg, p, flags):
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    cfg = So3kratesConfig(F=ocfg.F, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees), n_heads=ocfg.n_heads,
                          n_rbf=ocfg.n_rbf, r_cut=ocfg.r_cut, layer_normalization=ocfg.layer_normalization,
                          final_layer=ocfg.final_layer, residual_mlp_1=ocfg.residual_mlp_1,
                          residual_mlp_2=ocfg.residual_mlp_2)
    eng = So3kEngine(cfg, 'cuda:0', flags)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in p.items()})
    return eng


def dev(a, dt):
    return torch.as_tensor(np.asarray(a)).to('cuda:0', dt)


@pytest.mark.parametrize('flags', [0, 1, 3], ids=['fp32', 'tensor', 'tensor+keep'])
@pytest.mark.parametrize('name', list(OPTS))
def test_optional_layer_branches_ethanol(ethanol, name, flags):
    cfg = o.OracleConfig(F=132, n_layers=3, **OPTS[name])
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True, readout_gain=GAIN.get(name, 1.0))
    eng = make(cfg, p, flags)
    B = 2
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1)).copy()
    nl = o.get_indices(R, z, 5.0)
    R = np.concatenate([R, np.zeros((B, 1, 3))], 1)                      # one padding atom, three padding edge slots
    z = np.concatenate([z, np.zeros((B, 1), z.dtype)], 1)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 3), nl['idx_i'].dtype)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 3), nl['idx_j'].dtype)], 1)
    E, F, ea = eng.energy_forces(dev(R, torch.float32), dev(z, torch.int32), dev(ii, torch.int32), dev(jj, torch.int32),
                                 want_e_atom=True)
    assert eng.check_status() == 0
    E, F, ea = E.cpu().numpy().reshape(-1), F.cpu().numpy(), ea.cpu().numpy()
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, ii, jj)
    e_rtol, f_atol = tolerances(flags, float(np.abs(Fo.numpy()).max()))
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= e_rtol * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(F - Fo.numpy()).max() < f_atol
    assert (F[:, -1] == 0).all() and (ea[:, -1] == 0).all()


def test_optional_layer_branches_periodic_box():
    """A 600-atom periodic box (several blocks of every node kernel, ragged last block) with all branches on."""
    from mlff_b200.engine import So3kEngine  # noqa: F401  (fails loudly without the CUDA library)
    n, box = 600, 18.2
    pos = jittered_lattice(n, box, 5)
    z = water_like_numbers(n, 5)
    cell = np.eye(3) * box
    ii, jj, S = o.periodic_neighbor_list_kdtree(pos, box, 5.0)
    cfg = o.OracleConfig(F=132, n_layers=2, **OPTS['all'])
    p = o.init_params(cfg, 2, embed_scale=4.0, scale_shift=True, readout_gain=GAIN['all'])
    Eo, Fo, _ = o.energy_forces(cfg, p, pos, z, ii, jj, cell=cell, cell_offset=S)
    for flags in (0, 3):
        eng = make(cfg, p, flags)
        E, F, ea = eng.energy_forces(dev(pos[None], torch.float32), dev(z[None], torch.int32), dev(ii[None], torch.int32),
                                     dev(jj[None], torch.int32), dev(cell[None], torch.float32), dev(S[None], torch.int32),
                                     want_e_atom=True)
        assert eng.check_status() == 0
        e_rtol, f_atol = tolerances(flags, float(np.abs(Fo.numpy()).max()))
        dE, dF = abs(float(E.cpu()) - float(Eo)), np.abs(F.cpu().numpy()[0] - Fo.numpy()).max()
        print(f'periodic box, flags {flags}: |dE| {dE:.2e} (sum|e_atom| {float(ea.abs().sum()):.1f}), max|dF| {dF:.2e} '
              f'(max|F| {float(np.abs(Fo.numpy()).max()):.2f})')
        assert dE <= e_rtol * abs(float(Eo))
        assert dF < f_atol


def test_periodic_training_batch_through_get_pbc_neighbors(solid):
    """a padded batch of periodic frames of different sizes, cells and pbc goes through
    `get_pbc_neighbors` (mlff/indexing/indices.py:111-191) and the padding helpers straight into the batched
    energy+force call; every frame is checked against the oracle on its own ASE-style list."""
    from mlff_b200 import indexing as ix
    from common import as_sets
    rng = np.random.default_rng(0)
    R0, z0, cell0 = solid['R'], solid['z'], solid['unit_cell']
    frames = [(R0, z0, cell0, (1, 1, 1)), (R0[:70] + rng.normal(0, 0.05, (70, 3)), z0[:70], cell0 * 1.3, (1, 1, 0)),
              (R0[:25], z0[:25], cell0, (0, 0, 0))]
    n_max = max(len(f[0]) for f in frames)
    pos = np.concatenate([ix.pad_coordinates(f[0][None], n_max) for f in frames])
    z = np.concatenate([ix.pad_atomic_types(f[1][None], n_max) for f in frames])
    cells = np.stack([f[2] for f in frames])
    cfg = o.OracleConfig(F=132, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, 3)
    nb = ix.get_pbc_neighbors(eng, pos, z != 0, np.array([f[3] for f in frames]), cells, 5.0)
    assert eng.check_status() == 0
    E, F, ea = eng.energy_forces(dev(pos, torch.float32), dev(z, torch.int32), nb['idx_i'], nb['idx_j'],
                                 dev(cells, torch.float32), nb['shifts'], want_e_atom=True)
    assert eng.check_status() == 0
    for b, (R, zz, cell, pbc) in enumerate(frames):
        oi, oj, oS = o.primitive_neighbor_list(R, cell, pbc, 5.0)
        assert as_sets(*(nb[k][b].cpu().numpy() for k in ('idx_i', 'idx_j', 'shifts'))) == as_sets(oi, oj, oS)
        Eo, Fo, _ = o.energy_forces(cfg, p, R, zz, oi, oj, cell=cell, cell_offset=oS)
        assert abs(float(E[b].cpu()) - float(Eo)) <= E_RTOL * abs(float(Eo))
        assert np.abs(F[b, :len(R)].cpu().numpy() - Fo.numpy()).max() < F_ATOL
        assert float(F[b, len(R):].abs().sum()) == 0.0


def test_c5_wide_model_at_its_configured_size():
    """BASELINE.json configs[4] at its own size: 1 000 atoms in a 21.544 A periodic cell, F = 256, degrees 1-4 (M = 24),
    3 layers -- against the float64 oracle, in the fp32 CUDA-core mode and on the tensor cores.  F = 256 exceeds the narrow
    tcgen05 kernels' limit (weights resident in shared memory: F <= 144), so each filter block runs as 2 x 2 sub-blocks of
    them (so3k_edge_tc_split == 2).  A width that does not fit even halved must fail loudly (SO3K_ERR_CONFIG)."""
    from mlff_b200.capi import So3kError
    n, box = 1000, 21.544
    pos = jittered_lattice(n, box, 4)
    z = water_like_numbers(n, 4)
    cell = np.eye(3) * box
    ii, jj, S = o.periodic_neighbor_list_kdtree(pos, box, 5.0)
    cfg = o.OracleConfig(F=256, n_layers=3, degrees=(1, 2, 3, 4))
    p = o.init_params(cfg, 4, embed_scale=4.0, scale_shift=True)
    Eo, Fo, _ = o.energy_forces(cfg, p, pos, z, ii, jj, cell=cell, cell_offset=S)
    for flags in (0, 1, 3):
        eng = make(cfg, p, flags)
        E, F, _ = eng.energy_forces(dev(pos[None], torch.float32), dev(z[None], torch.int32), dev(ii[None], torch.int32),
                                    dev(jj[None], torch.int32), dev(cell[None], torch.float32), dev(S[None], torch.int32))
        assert eng.check_status() == 0
        dE, dF = abs(float(E.cpu()) - float(Eo)), np.abs(F.cpu().numpy()[0] - Fo.numpy()).max()
        print(f'c5 (1000 atoms, F=256, L_max=4, {len(ii)} edges), flags {flags}: |dE|/|E| {dE / abs(float(Eo)):.2e}, '
              f'max|dF| {dF:.2e} (max|F| {float(np.abs(Fo.numpy()).max()):.2f})')
        assert dE <= E_RTOL * abs(float(Eo))
        assert dF < F_ATOL
    wide = o.OracleConfig(F=320, n_layers=1, degrees=(1, 2, 3, 4))
    with pytest.raises(So3kError, match='SO3K_ERR_CONFIG'):
        make(wide, o.init_params(wide, 0), 1)
