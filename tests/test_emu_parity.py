"""The SIMT kernels, executed on the CPU through the emulation layer (tests/emu), against the oracle.
This exercises the same kernel SOURCE that nvcc compiles for sm_100a and the same C ABI, so indexing / masking /
graph-construction logic is checked without a GPU.  The `-m gpu` suite repeats the comparison on the real device."""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from emu_engine import EmuEngine, ptr, build_emu
from mlff_b200.capi import So3kLib, NL_ASE, NL_DENSE, NL_MIC, STATUS_ASYMMETRIC, STATUS_OVERFLOW
from common import as_sets, engine_flags, check_mic_list

E_RTOL = 1e-5      # north_star: energies within 1e-5 relative
F_ATOL = 1e-4      # north_star: forces within 1e-4 eV/A max-abs (fp32)


def make(cfg, p):
    eng = EmuEngine(cfg.F, cfg.n_rbf, cfg.n_layers, cfg.n_heads, cfg.degrees, cfg.r_cut,
                    flags=engine_flags(cfg), zbl_shift=cfg.zbl_repulsion_shift)
    eng.load(p)
    return eng


@pytest.mark.parametrize('F,degrees,H', [(36, (1, 2, 3), 4), (32, (1, 2, 2, 3), 4), (32, (1, 2, 3, 4), 2)])
def test_featurise_matches_oracle(ethanol, F, degrees, H):
    cfg = o.OracleConfig(F=F, degrees=degrees, n_heads=H, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    R = ethanol['R'][0]
    nl = o.get_indices(ethanol['R'][:1], ethanol['z'][None], 5.0)
    ii = np.concatenate([nl['idx_i'][0], [-1, -1, -1]])
    jj = np.concatenate([nl['idx_j'][0], [-1, -1, -1]])
    out = eng.featurise(9, ii, jj, R=R)
    geo = o.geometry_embed(cfg, torch.as_tensor(ii), torch.as_tensor(jj), torch.as_tensor((ii != -1).astype(np.float64)),
                           torch.float64, R=torch.as_tensor(R))
    np.testing.assert_allclose(out['d'], geo['d_ij'].numpy(), rtol=1e-6)
    np.testing.assert_allclose(out['unit'], geo['unit_r_ij'].numpy(), atol=1e-6)
    np.testing.assert_allclose(out['phi'], geo['phi_r_cut'].numpy(), atol=1e-6)
    np.testing.assert_allclose(out['rbf'], geo['rbf_ij'].numpy(), atol=2e-6)
    np.testing.assert_allclose(out['sph'], geo['sph_ij'].numpy(), atol=2e-6)
    assert (out['rbf'][-3:] == 0).all() and (out['sph'][-3:] == 0).all() and (out['phi'][-3:] == 0).all()


@pytest.mark.parametrize('F,L,degrees,H', [(132, 3, (1, 2, 3), 4), (36, 2, (1, 2, 3), 4), (32, 2, (1, 2, 2, 3), 4),
                                           (64, 2, (1, 2, 3, 4), 4)])
def test_energy_forces_match_oracle_ethanol(ethanol, F, L, degrees, H):
    cfg = o.OracleConfig(F=F, n_layers=L, degrees=degrees, n_heads=H)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p)
    B = 2
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    E, Fo, ea, st = eng.energy_forces(R, z, nl['idx_i'], nl['idx_j'])
    assert st == 0
    Eo, Fr = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(Fo - Fr.numpy()).max() < F_ATOL
    assert np.abs(Fo - Fr.numpy()).max() < 2e-5 * np.abs(Fr.numpy()).max()


def test_golden_default_model(ethanol):
    g = golden('oracle_ethanol_amplified.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    eng = make(cfg, p)
    z = np.tile(ethanol['z'][None], (4, 1))
    E, F, ea, st = eng.energy_forces(ethanol['R'][:4], z, g['idx_i'], g['idx_j'])
    np.testing.assert_allclose(E, g['E'].reshape(-1), rtol=E_RTOL)
    assert np.abs(F - g['F']).max() < F_ATOL


def test_padding_invariance_and_shuffled_edges(ethanol):
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p)
    R = ethanol['R'][:1]
    z = ethanol['z'][None]
    nl = o.get_indices(R, z, 5.0)
    E, F, _, _ = eng.energy_forces(R, z, nl['idx_i'], nl['idx_j'])
    # extra zero-Z atoms, extra -1 edges, and a shuffled edge order (the engine sorts by receiver itself)
    rng = np.random.default_rng(0)
    perm = rng.permutation(72)
    ii = np.concatenate([nl['idx_i'][0][perm][:30], [-1, -1], nl['idx_i'][0][perm][30:], [-1] * 3])[None]
    jj = np.concatenate([nl['idx_j'][0][perm][:30], [-1, -1], nl['idx_j'][0][perm][30:], [-1] * 3])[None]
    Rp = np.concatenate([R[0], np.zeros((3, 3))])[None]
    zp = np.concatenate([z[0], [0, 0, 0]])[None]
    Ep, Fp, _, st = eng.energy_forces(Rp, zp, ii, jj)
    assert st == 0
    np.testing.assert_allclose(Ep, E, rtol=2e-6)
    np.testing.assert_allclose(Fp[0, :9], F[0], atol=2e-6)
    assert (Fp[0, 9:] == 0).all()


def test_periodic_solid_with_cell_offsets(solid):
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    eng = make(cfg, p)
    E, F, ea, st, stress = eng.energy_forces(solid['R'][None], solid['z'][None], g['idx_i'][None], g['idx_j'][None],
                                              cell=solid['unit_cell'][None], cell_offset=g['shifts'][None],
                                              want_stress=True)
    assert st == 0
    assert abs(E[0] - float(g['E'])) <= E_RTOL * abs(float(g['E']))
    assert np.abs(F[0] - g['F']).max() < F_ATOL
    np.testing.assert_allclose(ea[0], g['e_atom'], atol=2e-5)
    # so3k_stress (k_stress): the strain derivative of the same evaluation (several periodic images per pair)
    _, _, So = o.energy_forces_stress(cfg, p, solid['R'], solid['z'], g['idx_i'], g['idx_j'], cell=solid['unit_cell'],
                                      cell_offset=g['shifts'])
    scale = max(np.abs(So.numpy()).max(), 1.0)
    assert np.abs(stress[0] - So.numpy()).max() < 1e-4 * scale
    assert np.abs(stress[0] - stress[0].T).max() < 1e-6 * scale


def test_md_seam_displacements(ethanol):
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p)
    R = ethanol['R'][0]
    z = ethanol['z']
    nl = o.get_indices(R[None], z[None], 5.0)
    ii, jj = nl['idx_i'][0], nl['idx_j'][0]
    N = len(z)
    # glp conventions: sentinel index N and a boolean mask
    centers = np.concatenate([ii, [N, N]])
    others = np.concatenate([jj, [N, N]])
    edges = np.concatenate([R[jj] - R[ii], np.zeros((2, 3))])
    mask = np.concatenate([np.ones(72, bool), [False, False]])
    ea, dE, Fo, st = eng.potential(edges, z, centers, others, mask, energy_scale=1.7)
    assert st == 0
    ea_o, g_o = o.potential_displacements(cfg, p, edges, z, np.where(mask, centers, -1), np.where(mask, others, -1),
                                          mask, energy_scale=1.7)
    np.testing.assert_allclose(ea, ea_o.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dE, g_o.numpy(), rtol=2e-5, atol=5e-6)
    _, F_o, _ = o.energy_forces(cfg, p, R, z, ii, jj)
    np.testing.assert_allclose(Fo, 1.7 * F_o.numpy(), rtol=2e-5, atol=1e-5)


def test_asymmetric_edge_list_is_flagged(ethanol):
    cfg = o.OracleConfig(F=36, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    R = ethanol['R'][:1]
    z = ethanol['z'][None]
    nl = o.get_indices(R, z, 5.0)
    ii, jj = nl['idx_i'].copy(), nl['idx_j'].copy()
    ii[0, 5] = -1
    jj[0, 5] = -1          # drop one direction of a pair
    _, _, _, st = eng.energy_forces(R, z, ii, jj)
    assert st & STATUS_ASYMMETRIC


# ---- neighbour list ---------------------------------------------------------------------------------
def run_nl(lib, mode, pos, z, cell, pbc, cutoff, capacity):
    N = len(pos)
    pos = np.ascontiguousarray(pos, np.float64)
    zz = None if z is None else np.ascontiguousarray(z, np.int32)
    cell = None if cell is None else np.ascontiguousarray(cell, np.float64)
    pbc = np.ascontiguousarray(pbc, np.uint8)
    ii = np.zeros(capacity, np.int32)
    jj = np.zeros(capacity, np.int32)
    S = np.zeros((capacity, 3), np.int32)
    cnt = np.zeros(1, np.int64)
    st = np.zeros(1, np.int32)
    nb = lib.neighbor_list_workspace_bytes(N, capacity)
    ws = np.zeros(nb // 8 + 8, np.float64)
    lib.neighbor_list(mode, N, ptr(pos), ptr(zz), ptr(cell), ptr(pbc), cutoff, capacity, ptr(ii), ptr(jj), ptr(S),
                      ptr(cnt), ptr(ws), ws.nbytes, ptr(st))
    return ii, jj, S, int(cnt[0]), int(st[0])


@pytest.mark.parametrize('pbc', [(1, 1, 1), (1, 1, 0)])
def test_minimum_image_list_matches_glp_semantics(solid, pbc):
    """the MD path's list (glp quadratic_neighbor_list, mlff/mdx/atoms.py:488-529,
    mlff/md/calculator.py:204-238) -- device mode SO3K_NL_MIC against oracle.mic_neighbor_list, on a cell SMALLER than
    2 r_cut (9.05 A: the all-image list has several images per pair, the minimum-image list exactly one), on a skewed
    cell, and fed through the MD seam (displacements in, per-atom energies + dE/d(edges) out)."""
    lib = So3kLib(build_emu())
    R, cell = solid['R'], solid['unit_cell']
    n_all = len(o.primitive_neighbor_list(R, cell, pbc, 5.0)[0])
    ii, jj, S, cnt, st = run_nl(lib, NL_MIC, R, None, cell, pbc, 5.0, n_all)
    assert st == 0
    oi, oj, odv = check_mic_list(o, R, cell, pbc, 5.0, ii, jj, S, cnt)
    assert cnt < n_all                                                 # fewer edges than the all-image list
    rng = np.random.default_rng(1)
    cell3 = np.array([[7.0, 0, 0], [2.0, 6.5, 0], [1.0, -1.5, 8.0]])
    p3 = rng.uniform(-1, 2, (150, 3)) @ cell3
    ii, jj, S, cnt, st = run_nl(lib, NL_MIC, p3, None, cell3, (1, 1, 1), 3.0, 20000)
    assert st == 0
    check_mic_list(o, p3, cell3, (1, 1, 1), 3.0, ii, jj, S, cnt)
    if pbc == (1, 1, 1):
        # ... and through the MD seam, as CalculatorX / mlffCalculator use it: edges = minimum-image displacements
        cfg = o.OracleConfig(F=36, n_layers=2)
        p = o.init_params(cfg, 0, embed_scale=4.0)
        ea, dE, Fo, st = make(cfg, p).potential(odv, solid['z'], oi, oj)
        ea_o, dE_o = o.potential_displacements(cfg, p, odv, solid['z'], oi, oj, np.ones(len(oi), bool))
        assert st == 0
        np.testing.assert_allclose(ea, ea_o.numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(dE, dE_o.numpy(), rtol=5e-5, atol=5e-5)


@pytest.mark.parametrize('pbc', [(1, 1, 1), (1, 1, 0), (0, 0, 0), (1, 0, 1)])
def test_neighbor_list_bit_exact_small_cell(solid, pbc):
    # 9.05 A cell < 2 * 5 A cutoff: several periodic images of the same pair
    lib = So3kLib(build_emu())
    oi, oj, oS = o.primitive_neighbor_list(solid['R'], solid['unit_cell'], pbc, 5.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_ASE, solid['R'], None, solid['unit_cell'], pbc, 5.0, len(oi) + 5)
    assert st == 0 and cnt == len(oi)
    assert (ii[:cnt] == oi).all() and (jj[:cnt] == oj).all() and (S[:cnt] == oS).all()
    assert (ii[cnt:] == -1).all() and (jj[cnt:] == -1).all() and (S[cnt:] == 0).all()      # indices.py:226-231


def test_neighbor_list_overflow_triclinic_unwrapped_dense(solid, ethanol):
    lib = So3kLib(build_emu())
    oi, oj, oS = o.primitive_neighbor_list(solid['R'], solid['unit_cell'], (1, 1, 1), 5.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_ASE, solid['R'], None, solid['unit_cell'], (1, 1, 1), 5.0, 100)
    assert (st & STATUS_OVERFLOW) and cnt == len(oi)                                        # indices.py:264-272
    # atoms far outside the cell (ASE does not wrap; shifts absorb it)
    pos2 = solid['R'] + np.array([13.0, -22.0, 40.0])
    oi, oj, oS = o.primitive_neighbor_list(pos2, solid['unit_cell'], (1, 1, 1), 5.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_ASE, pos2, None, solid['unit_cell'], (1, 1, 1), 5.0, len(oi))
    assert cnt == len(oi) and (ii == oi).all() and (jj == oj).all() and (S == oS).all()
    # triclinic
    rng = np.random.default_rng(0)
    cell3 = np.array([[7.0, 0, 0], [2.0, 6.5, 0], [1.0, -1.5, 8.0]])
    p3 = rng.uniform(0, 1, (60, 3)) @ cell3
    oi, oj, oS = o.primitive_neighbor_list(p3, cell3, (1, 1, 1), 4.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_ASE, p3, None, cell3, (1, 1, 1), 4.0, len(oi) + 3)
    assert cnt == len(oi) and (ii[:cnt] == oi).all() and (jj[:cnt] == oj).all() and (S[:cnt] == oS).all()
    # dense (training) list: 0 < d <= r_cut, row-major, zero-Z atoms excluded
    R = ethanol['R'][0]
    z = ethanol['z'].copy()
    nl = o.get_indices(R[None], z[None], 5.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_DENSE, R, z, None, (0, 0, 0), 5.0, 80)
    assert cnt == 72 and (ii[:72] == nl['idx_i'][0]).all() and (jj[:72] == nl['idx_j'][0]).all()
    z[4] = 0
    nl = o.get_indices(R[None], z[None], 5.0)
    ii, jj, S, cnt, st = run_nl(lib, NL_DENSE, R, z, None, (0, 0, 0), 5.0, 80)
    k = (nl['idx_i'][0] >= 0).sum()
    assert cnt == k and as_sets(ii, jj) == as_sets(nl['idx_i'][0], nl['idx_j'][0])


def test_zbl_repulsion(ethanol):
    # Energy(zbl_repulsion=True), observable.py:81-84, 110-183: both seams, shift conventions included
    cfg = o.OracleConfig(F=36, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.37)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p)
    R = ethanol['R'][:2]
    z = np.tile(ethanol['z'][None], (2, 1))
    nl = o.get_indices(R, z, 5.0)
    E, F, ea, st = eng.energy_forces(R, z, nl['idx_i'], nl['idx_j'])
    assert st == 0
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    E0, F0 = o.energy_forces_batch(o.OracleConfig(F=36, n_layers=2), p, R, z, nl['idx_i'], nl['idx_j'])
    assert np.abs(Fo.numpy() - F0.numpy()).max() > 1.0          # the repulsion matters here (C-H, O-H bonds < 1.5 A)
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(F - Fo.numpy()).max() < F_ATOL
    ii, jj = nl['idx_i'][0], nl['idx_j'][0]
    edges = R[0][jj] - R[0][ii]
    mask = np.ones(len(ii), bool)
    out = eng.potential(edges, z[0], ii, jj, mask, energy_scale=1.3)
    ea_o, dE_o = o.potential_displacements(cfg, p, edges, z[0], ii, jj, mask, energy_scale=1.3)
    np.testing.assert_allclose(out[0], ea_o.numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(out[1], dE_o.numpy(), rtol=2e-5, atol=2e-5)


def test_stress_of_a_molecular_batch(ethanol):
    # so3k_stress: one tensor per structure of a padded batch (the periodic case is part of
    # test_periodic_solid_with_cell_offsets)
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 3, embed_scale=4.0)
    eng = make(cfg, p)
    B = 2
    Rb = ethanol['R'][:B]
    zb = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(Rb, zb, 5.0)
    _, _, _, st, stb = eng.energy_forces(Rb, zb, nl['idx_i'], nl['idx_j'], want_stress=True)
    assert st == 0
    for b in range(B):
        _, _, So = o.energy_forces_stress(cfg, p, Rb[b], zb[b], nl['idx_i'][b], nl['idx_j'][b])
        assert np.abs(stb[b] - So.numpy()).max() < 1e-4 * max(np.abs(So.numpy()).max(), 1.0)


@pytest.mark.parametrize('opts', [dict(layer_normalization=True), dict(layer_normalization=True, final_layer=True),
                                  dict(residual_mlp_1=True), dict(residual_mlp_2=True),
                                  dict(layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True)],
                         ids=['ln', 'ln+post', 'res1', 'res2', 'all'])
def test_optional_layer_branches(ethanol, opts):
    # So3kratesLayer(layer_normalization / final_layer / residual_mlp_1 / residual_mlp_2), so3krates_layer.py:103-174
    cfg = o.OracleConfig(F=36, n_layers=2, **opts)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p)
    B = 2
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1)).copy()
    nl = o.get_indices(R, z, 5.0)
    # one padding atom and padding edge slots: the masked LayerNorm rows must stay unobservable
    R = np.concatenate([R, np.zeros((B, 1, 3))], 1)
    z = np.concatenate([z, np.zeros((B, 1), z.dtype)], 1)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 3), nl['idx_i'].dtype)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 3), nl['idx_j'].dtype)], 1)
    E, Fo, ea, st = eng.energy_forces(R, z, ii, jj)
    assert st == 0
    Eo, Fr = o.energy_forces_batch(cfg, p, R, z, ii, jj)
    base = o.OracleConfig(F=36, n_layers=2)
    E0, F0 = o.energy_forces_batch(base, p, R, z, ii, jj)
    assert np.abs(Fr.numpy() - F0.numpy()).max() > 1e-2         # the branch changes the model
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(Fo - Fr.numpy()).max() < F_ATOL
    assert np.abs(Fo - Fr.numpy()).max() < 5e-5 * np.abs(Fr.numpy()).max()
    assert (Fo[:, -1] == 0).all() and (ea[:, -1] == 0).all()
    if len(opts) == 4:                                          # the MD seam shares the stages: displacements in, dE/d(edges) out
        i0, j0 = nl['idx_i'][0], nl['idx_j'][0]
        edges = ethanol['R'][0][j0] - ethanol['R'][0][i0]
        mask = np.ones(len(i0), bool)
        out = eng.potential(edges, ethanol['z'], i0, j0, mask, energy_scale=0.7)
        ea_o, dE_o = o.potential_displacements(cfg, p, edges, ethanol['z'], i0, j0, mask, energy_scale=0.7)
        np.testing.assert_allclose(out[0], ea_o.numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(out[1], dE_o.numpy(), rtol=5e-5, atol=5e-5)


class EmuNLEngine:
    """`engine.neighbor_list` of mlff_b200.engine.So3kEngine, backed by the emulated cell-list kernels (host tensors)."""
    device = torch.device('cpu')

    def __init__(self):
        self.lib = So3kLib(build_emu())
        self._status = torch.zeros(1, dtype=torch.int32)

    def neighbor_list(self, positions, cutoff, capacity, cell=None, pbc=(False, False, False), z=None, mode=NL_ASE):
        zz = None if z is None else z.numpy()
        ii, jj, S, cnt, st = run_nl(self.lib, mode, positions.numpy(), zz, cell, pbc, cutoff, capacity)
        self._status |= st
        return torch.as_tensor(ii), torch.as_tensor(jj), torch.as_tensor(S), torch.tensor([cnt])


def test_get_pbc_neighbors_and_padding_helpers(solid):
    # mlff/indexing/indices.py:111-191 -- a padded batch of periodic frames with different sizes, cells and pbc
    from mlff_b200 import indexing as ix
    rng = np.random.default_rng(0)
    R0, cell0 = solid['R'], solid['unit_cell']
    frames = [(R0, cell0, (1, 1, 1)), (R0[:70] + rng.normal(0, 0.05, (70, 3)), cell0 * 1.3, (1, 1, 0)),
              (R0[:25], cell0, (0, 0, 0))]
    n_max = max(len(f[0]) for f in frames)
    pos = np.concatenate([ix.pad_coordinates(f[0][None], n_max) for f in frames])
    mask = np.stack([np.arange(n_max) < len(f[0]) for f in frames])
    out = ix.get_pbc_neighbors(EmuNLEngine(), pos, mask, np.array([f[2] for f in frames]),
                               np.stack([f[1] for f in frames]), 5.0)
    lens = []
    for b, (R, cell, pbc) in enumerate(frames):
        oi, oj, oS = o.primitive_neighbor_list(R, cell, pbc, 5.0)
        lens.append(len(oi))
        ii, jj, S = (out[k][b].numpy() for k in ('idx_i', 'idx_j', 'shifts'))
        assert as_sets(ii, jj, S) == as_sets(oi, oj, oS)
        assert (ii[len(oi):] == -1).all() and (jj[len(oi):] == -1).all() and (S[len(oi):] == 0).all()
        assert ii[:len(oi)].max() < len(R)                      # padding atoms never appear
    assert out['idx_i'].shape == (3, max(lens)) and out['shifts'].shape == (3, max(lens), 3)
    # ... and straight into the batched energy+force call (per-frame cells, padding atoms and slots)
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    zs = [solid['z'][:len(f[0])] for f in frames]
    z = np.concatenate([ix.pad_atomic_types(zz[None], n_max) for zz in zs])
    E, F, ea, st = make(cfg, p).energy_forces(pos, z, out['idx_i'].numpy(), out['idx_j'].numpy(),
                                              np.stack([f[1] for f in frames]), out['shifts'].numpy())
    assert st == 0
    for b, (R, cell, pbc) in enumerate(frames):
        oi, oj, oS = o.primitive_neighbor_list(R, cell, pbc, 5.0)
        Eo, Fo, _ = o.energy_forces(cfg, p, R, zs[b], oi, oj, cell=cell, cell_offset=oS)
        assert abs(E[b] - float(Eo)) <= E_RTOL * abs(float(Eo))
        assert np.abs(F[b, :len(R)] - Fo.numpy()).max() < F_ATOL and (F[b, len(R):] == 0).all()
    holed = mask.copy()
    holed[1, 3] = False                                         # a padding atom in front of real ones
    with pytest.raises(ValueError):
        ix.get_pbc_neighbors(EmuNLEngine(), pos, holed, np.ones((3, 3), bool), np.stack([f[1] for f in frames]), 5.0)
    empty = ix.get_pbc_neighbors(EmuNLEngine(), pos[:1], np.zeros((1, n_max), bool), np.ones((1, 3), bool), cell0[None], 5.0)
    assert empty['idx_i'].shape == (1, 0)
    # padding helpers: numpy and torch, the reference's shapes and fill values (mlff/padding/padding.py:47-108)
    z = np.arange(1, 7).reshape(2, 3)
    assert ix.pad_atomic_types(z, 5).tolist() == [[1, 2, 3, 0, 0], [4, 5, 6, 0, 0]]
    assert ix.pad_coordinates(torch.ones(2, 3, 3), 4).shape == (2, 4, 3) and ix.pad_forces(np.ones((2, 3, 3)), 4)[:, 3].sum() == 0
    pi, pj = ix.pad_indices(np.zeros((2, 4), np.int32), torch.zeros(2, 4, dtype=torch.int32), 6)
    assert pi.shape == (2, 6) and (pi[:, 4:] == -1).all() and bool((pj[:, 4:] == -1).all())
    assert ix.pad_index_list(np.array([1, 2]), 4).tolist() == [1, 2, -1, -1]
    assert ix.pad_shift(np.ones((2, 3), np.int32), 3).tolist() == [[1, 1, 1], [1, 1, 1], [0, 0, 0]]
    assert ix.pad_per_atom_quantity(np.ones((1, 2, 1)), 3).shape == (1, 3, 1)
    with pytest.raises(AssertionError):
        ix.pad_atomic_types(z, 2)
    # ... and against the outputs of the reference's own padding functions (tests/golden/reference_functions.npz)
    g = golden('reference_functions.npz')
    zz, RR, ij = g['pad/in/z'], g['pad/in/R'], g['pad/in/idx']
    assert np.array_equal(ix.pad_atomic_types(zz, 5), g['pad/atomic_types'])
    assert np.array_equal(ix.pad_coordinates(RR, 5), g['pad/coordinates']) and np.array_equal(ix.pad_forces(RR, 4), g['pad/forces'])
    assert np.array_equal(ix.pad_per_atom_quantity(RR[..., :1], 6), g['pad/per_atom_quantity'])
    pi, pj = ix.pad_indices(ij, ij[:, ::-1], 7)
    assert np.array_equal(pi, g['pad/indices_i']) and np.array_equal(pj, g['pad/indices_j'])
    assert np.array_equal(ix.pad_index_list(ij[0], 6), g['pad/index_list']) and np.array_equal(ix.pad_shift(ij[:, :3], 5), g['pad/shift'])


def test_device_dense_list_matches_reference_get_indices():
    """The cell-list kernel in SO3K_NL_DENSE mode (emulated) against arrays produced by the reference's own get_indices
    source (tests/golden/reference_functions.npz): identical (i, j) sequences per frame -- the kernel's order (sorted by
    i, then j) is the reference's row-major order."""
    g = golden('reference_functions.npz')
    lib = So3kLib(build_emu())
    R, z = g['get_indices/R'], g['get_indices/z']
    for rc in (5.0, 2.0):
        ref_i, ref_j = g[f'get_indices/{rc}/idx_i'], g[f'get_indices/{rc}/idx_j']
        for b in range(len(R)):
            ii, jj, S, cnt, st = run_nl(lib, NL_DENSE, R[b], z[b], None, (0, 0, 0), rc, 128)
            k = int((ref_i[b] >= 0).sum())
            assert st == 0 and cnt == k
            assert np.array_equal(ii[:k], ref_i[b][:k]) and np.array_equal(jj[:k], ref_j[b][:k])
            assert (ii[k:] == -1).all() and (S == 0).all()


REFERENCE_MODEL_CASES = {
    'ethanol_default': (dict(), dict(scale_shift=True)),
    'ethanol_1223': (dict(F=32, n_layers=2, degrees=(1, 2, 2, 3), n_heads=2), {}),
    'ethanol_all_branches': (dict(F=36, n_layers=2, layer_normalization=True, final_layer=True, residual_mlp_1=True,
                                  residual_mlp_2=True), {}),
    'ethanol_zbl': (dict(F=36, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.37), dict(scale_shift=True)),
    'solid_periodic': (dict(F=36, n_layers=2), {})}


@pytest.mark.parametrize('tag', list(REFERENCE_MODEL_CASES))
def test_kernels_match_reference_model_code(tag):
    """The kernels (emulated) DIRECTLY against energies and finite-difference forces produced by the reference's own
    model code (tests/golden/reference_model.npz, see tests/golden/make_reference_model_golden.py) -- no oracle in
    between: default model with padding, repeated degrees, all optional layer branches, ZBL, periodic solid."""
    g = golden('reference_model.npz')
    ckw, pkw = REFERENCE_MODEL_CASES[tag]
    cfg = o.OracleConfig(**ckw)
    p = o.init_params(cfg, int(g[f'{tag}/seed']), embed_scale=4.0, **pkw)      # the generator's parameters
    ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith(tag + '/in/')}
    cell = ins['unit_cell'][None] if 'unit_cell' in ins else None
    off = ins['cell_offset'][None] if 'cell_offset' in ins else None
    E, F, ea, st = make(cfg, p).energy_forces(ins['R'][None], ins['z'][None], ins['idx_i'][None], ins['idx_j'][None], cell, off)
    assert st == 0
    E_ref, F_fd, atoms = float(g[f'{tag}/E']), g[f'{tag}/F_fd'], g[f'{tag}/fd_atoms']
    assert abs(float(E[0]) - E_ref) <= E_RTOL * abs(E_ref)
    assert np.abs(F[0][atoms] - F_fd).max() < F_ATOL


@pytest.mark.parametrize('keep', [0, 2])
@pytest.mark.parametrize('F,degrees,H', [(132, (1, 2, 3), 4), (36, (1, 2, 3), 4), (32, (1, 2, 2, 3), 4), (144, (1, 2, 3), 8),
                                         (140, (1, 2, 3, 4), 4), (256, (1, 2, 3, 4), 4), (160, (1, 2, 3, 4), 4)])
def test_tensor_mode_orchestration_and_streaming_kernels(ethanol, solid, F, degrees, H, keep):
    """Tensor mode on the emulator: the two tcgen05 pair kernels are stood in for by plain fp32 loops (so3k_edge_tc.cu,
    SO3K_EMU branch); everything around them is the product source -- pair records, canonical pairs, the streaming
    kernels that consume pair-indexed filters and emit pair-indexed filter cotangents (k_alpha_aggregate -- the warp-scan
    variant for F = 132 / 36 / 32, the general one for F = 140 / 144 / 256 --, k_alpha_bwd, k_qk_bwd_stream), kept /
    recomputed filters.  F = 160 and 256 exceed the narrow kernels' limit (the stand-ins keep it): they run as 2 x 2
    sub-blocks -- sliced weights, sub-block outputs combined into the F-wide rows, cotangents accumulated over the
    sub-blocks -- which is the host logic the B200 uses for the wide model (BASELINE configs[4])."""
    from mlff_b200.capi import FLAG_TENSOR
    cfg = o.OracleConfig(F=F, n_layers=2, degrees=degrees, n_heads=H)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = EmuEngine(cfg.F, cfg.n_rbf, cfg.n_layers, cfg.n_heads, cfg.degrees, cfg.r_cut, flags=FLAG_TENSOR | keep)
    eng.load(p)
    B = 2
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    # padding atoms and padding pair slots, as the training batches have them
    R = np.concatenate([R, np.zeros((B, 2, 3))], 1)
    z = np.concatenate([z, np.zeros((B, 2), z.dtype)], 1)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 5), nl['idx_i'].dtype)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 5), nl['idx_j'].dtype)], 1)
    E, Fc, ea, st = eng.energy_forces(R, z, ii, jj)
    assert st == 0
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, ii, jj)
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(Fc - Fo.numpy()).max() < F_ATOL and (Fc[:, -2:] == 0).all()
    if F == 36:
        # a periodic solid with several images per pair (cell < 2 r_cut): long receiver rows (> 32 edges per chunk)
        Rs, zs, cell = solid['R'], solid['z'], solid['unit_cell']
        oi, oj, oS = o.primitive_neighbor_list(Rs, cell, (True, True, True), 5.0)
        E, Fc, ea, st = eng.energy_forces(Rs[None], zs[None], oi[None], oj[None], cell[None], oS[None])
        Eo, Fo, _ = o.energy_forces(cfg, p, Rs, zs, oi, oj, cell=cell, cell_offset=oS)
        assert st == 0 and abs(E[0] - float(Eo)) <= E_RTOL * abs(float(Eo))
        assert np.abs(Fc[0] - Fo.numpy()).max() < F_ATOL
