"""Parity of the CUDA path (through the C ABI, libso3k.so) against the CPU oracle and the committed golden vectors.
Run on the B200 box:  python -m pytest tests -m gpu -x -q
Tolerances are north_star's: neighbour lists bit-exact as sorted (i,j,S) sets, energies <= 1e-5 relative,
forces <= 1e-4 eV/A max-abs (fp32)."""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from common import as_sets, jittered_lattice, water_like_numbers

pytestmark = pytest.mark.gpu

E_RTOL = 1e-5
F_ATOL = 1e-4


def _cfgs(ocfg):
    from mlff_b200.config import So3kratesConfig
    return So3kratesConfig(F=ocfg.F, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees), n_heads=ocfg.n_heads,
                           n_rbf=ocfg.n_rbf, r_cut=ocfg.r_cut, zbl_repulsion=ocfg.zbl_repulsion,
                           zbl_repulsion_shift=ocfg.zbl_repulsion_shift)


def make(ocfg, p, flags=0):
    from mlff_b200.engine import So3kEngine
    eng = So3kEngine(_cfgs(ocfg), 'cuda:0', flags)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in p.items()})
    return eng


def dev(a, dt):
    return torch.as_tensor(np.asarray(a)).to('cuda:0', dt)


def run_ef(eng, R, z, ii, jj, cell=None, off=None):
    E, F, ea = eng.energy_forces(dev(R, torch.float32), dev(z, torch.int32), dev(ii, torch.int32), dev(jj, torch.int32),
                                 None if cell is None else dev(cell, torch.float32),
                                 None if off is None else dev(off, torch.int32), want_e_atom=True)
    st = eng.check_status()
    return E.cpu().numpy().reshape(-1), F.cpu().numpy(), ea.cpu().numpy(), st


def check_md_seam_outputs(flags, ii, jj, ea_g, dE_g, F_g, ea_o, dE_o, n):
    """Outputs of so3k_potential_displacements against the oracle.  fp32 mode: per-edge dE/d(edges) equal.  Tensor mode
    evaluates the radial / l0 cotangents once per undirected pair and books them on one of its two directed edges, so
    dE/d(edges) equals the reference up to that antisymmetric pair gauge: every quantity the MD caller derives from it
    -- the pair difference dE_e - dE_rev(e), hence the forces (scatter over centers / others) and the stress -- is
    checked instead (include/so3k.h)."""
    ea_g, dE_g, F_g = ea_g.cpu().numpy(), dE_g.cpu().numpy(), F_g.cpu().numpy()
    tol = 2e-5 if flags == 0 else F_ATOL                   # split-bf16 filters: north_star's force tolerance
    np.testing.assert_allclose(ea_g, ea_o, rtol=2e-5, atol=tol)
    if flags == 0:
        np.testing.assert_allclose(dE_g, dE_o, rtol=2e-5, atol=tol)
    rev = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(ii, jj))}
    r = np.array([rev[(int(b), int(a))] for a, b in zip(ii, jj)])
    np.testing.assert_allclose(dE_g - dE_g[r], dE_o - dE_o[r], rtol=2e-5, atol=2 * tol)
    F_o = np.zeros((n, 3))
    np.add.at(F_o, ii, dE_o)                               # F_i = -dE/dR_i, edges = R_j - R_i
    np.add.at(F_o, jj, -dE_o)
    np.testing.assert_allclose(F_g, F_o, rtol=2e-5, atol=2 * tol)


def test_native_library_is_loaded():
    from mlff_b200 import _lib
    lib = _lib.load()
    assert lib.path.endswith('mlff_b200/lib/libso3k.so')
    with open('/proc/self/maps') as f:
        assert 'libso3k.so' in f.read()


@pytest.mark.parametrize('F,degrees,H', [(132, (1, 2, 3), 4), (32, (1, 2, 2, 3), 4), (256, (1, 2, 3, 4), 4)])
def test_featurise_matches_oracle(ethanol, F, degrees, H):
    cfg = o.OracleConfig(F=F, degrees=degrees, n_heads=H, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    R = ethanol['R'][0]
    nl = o.get_indices(ethanol['R'][:1], ethanol['z'][None], 5.0)
    ii = np.concatenate([nl['idx_i'][0], [-1, -1, -1]])
    jj = np.concatenate([nl['idx_j'][0], [-1, -1, -1]])
    out = {k: v.cpu().numpy() for k, v in eng.featurise(dev(ii, torch.int32), dev(jj, torch.int32), 9,
                                                        R=dev(R, torch.float32)).items()}
    geo = o.geometry_embed(cfg, torch.as_tensor(ii), torch.as_tensor(jj), torch.as_tensor((ii != -1).astype(np.float64)),
                           torch.float64, R=torch.as_tensor(R))
    for k, tol in (('d_ij', 2e-6), ('unit_r_ij', 1e-6), ('phi_r_cut', 1e-6), ('rbf_ij', 2e-6), ('sph_ij', 2e-6)):
        np.testing.assert_allclose(out[k], geo[k].numpy(), atol=tol, err_msg=k)
    assert (out['rbf_ij'][-3:] == 0).all() and (out['sph_ij'][-3:] == 0).all()


@pytest.mark.parametrize('F,L,degrees,H', [(132, 3, (1, 2, 3), 4), (36, 2, (1, 2, 3), 4), (32, 2, (1, 2, 2, 3), 4),
                                           (256, 3, (1, 2, 3, 4), 4)])
def test_energy_forces_match_oracle_ethanol_batch(ethanol, F, L, degrees, H):
    # config 1 of BASELINE.json: 32 x 9-atom ethanol frames, default model
    cfg = o.OracleConfig(F=F, n_layers=L, degrees=degrees, n_heads=H)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p)
    B = 32 if F == 132 else 4
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    E, Fo, ea, st = run_ef(eng, R, z, nl['idx_i'], nl['idx_j'])
    assert st == 0
    Eo, Fr = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(Fo - Fr.numpy()).max() < F_ATOL
    assert np.abs(Fo - Fr.numpy()).max() < 3e-5 * np.abs(Fr.numpy()).max()


def test_golden_vectors(ethanol, solid):
    cfg = o.OracleConfig()
    for tag in ('default', 'amplified'):
        g = golden(f'oracle_ethanol_{tag}.npz')
        eng = make(cfg, o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale'])))
        z = np.tile(ethanol['z'][None], (4, 1))
        E, F, ea, st = run_ef(eng, ethanol['R'][:4], z, g['idx_i'], g['idx_j'])
        np.testing.assert_allclose(E, g['E'].reshape(-1), rtol=E_RTOL)
        assert np.abs(F - g['F']).max() < F_ATOL
    g = golden('oracle_solid_pbc.npz')
    eng = make(cfg, o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale'])))
    E, F, ea, st = run_ef(eng, solid['R'][None], solid['z'][None], g['idx_i'][None], g['idx_j'][None],
                          cell=solid['unit_cell'][None], off=g['shifts'][None])
    assert st == 0
    assert abs(E[0] - float(g['E'])) <= E_RTOL * abs(float(g['E']))
    assert np.abs(F[0] - g['F']).max() < F_ATOL
    np.testing.assert_allclose(ea[0], g['e_atom'], atol=2e-5)


@pytest.mark.parametrize('flags', [0, 1, 3])
def test_rotation_and_padding_invariance(ethanol, flags):
    # reference tests/test_model_invariance.py:80-222 re-expressed on the CUDA path
    from scipy.spatial.transform import Rotation
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=flags)
    tight = flags == 0           # the split-bf16 filters are rotation invariant to their own rounding (~1e-6 relative)
    R = ethanol['R'][:2]
    z = np.tile(ethanol['z'][None], (2, 1))
    nl = o.get_indices(R, z, 5.0)
    E, F, _, _ = run_ef(eng, R, z, nl['idx_i'], nl['idx_j'])
    for k in range(3):
        Q = Rotation.random(random_state=k).as_matrix()
        Er, Fr, _, _ = run_ef(eng, R @ Q.T, z, nl['idx_i'], nl['idx_j'])
        assert np.isclose(Er, E, rtol=1e-5 if tight else 3e-5).all()
        assert np.isclose((F @ Q.T).reshape(-1), Fr.reshape(-1), atol=1e-5 if tight else 5e-5).all()
    Rp = np.concatenate([R, np.zeros((2, 3, 3))], 1)
    zp = np.concatenate([z, np.zeros((2, 3), z.dtype)], 1)
    iip = np.concatenate([nl['idx_i'], -np.ones((2, 5), np.int64)], 1)
    jjp = np.concatenate([nl['idx_j'], -np.ones((2, 5), np.int64)], 1)
    Ep, Fp, _, _ = run_ef(eng, Rp, zp, iip, jjp)
    np.testing.assert_allclose(Ep, E, rtol=2e-6)
    np.testing.assert_allclose(Fp[:, :9], F, atol=2e-6)
    assert (Fp[:, 9:] == 0).all()
    # ragged input: one molecule loses every edge of its last atom (an isolated atom), the other is untouched
    drop = (nl['idx_i'][0] == 8) | (nl['idx_j'][0] == 8)
    iir, jjr = nl['idx_i'].copy(), nl['idx_j'].copy()
    iir[0, drop] = -1
    jjr[0, drop] = -1
    Eg, Fg, _, st = run_ef(eng, R, z, iir, jjr)
    assert st == 0
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, iir, jjr)
    np.testing.assert_allclose(Eg, Eo.numpy().reshape(-1), rtol=1e-5)
    assert np.abs(Fg - Fo.numpy()).max() < F_ATOL
    assert np.abs(Fg[0, 8]).max() == 0.0
    # no edges at all
    En, Fn, _, st = run_ef(eng, R, z, -np.ones_like(iir), -np.ones_like(jjr))
    assert st == 0 and np.isfinite(En).all() and (Fn == 0).all()


@pytest.mark.parametrize('seam_flags', [0, 3])
def test_md_seam_and_asymmetric_flag(ethanol, seam_flags):
    from mlff_b200.capi import So3kError
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=seam_flags)
    R, z = ethanol['R'][0], ethanol['z']
    nl = o.get_indices(R[None], z[None], 5.0)
    ii, jj = nl['idx_i'][0], nl['idx_j'][0]
    N = len(z)
    centers = np.concatenate([ii, [N, N]])
    others = np.concatenate([jj, [N, N]])
    edges = np.concatenate([R[jj] - R[ii], np.zeros((2, 3))])
    mask = np.concatenate([np.ones(72, bool), [False, False]])
    ea, dE, Fo = eng.potential(dev(edges, torch.float32), dev(z, torch.int32), dev(centers, torch.int32),
                               dev(others, torch.int32), dev(mask, torch.uint8), energy_scale=1.7)
    assert eng.check_status() == 0
    ea_o, g_o = o.potential_displacements(cfg, p, edges, z, np.where(mask, centers, -1), np.where(mask, others, -1),
                                          mask, energy_scale=1.7)
    if seam_flags == 0:
        np.testing.assert_allclose(ea.cpu().numpy(), ea_o.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dE.cpu().numpy(), g_o.numpy(), rtol=2e-5, atol=5e-6)
    check_md_seam_outputs(seam_flags, ii, jj, ea[:N], dE[:72], Fo, ea_o.numpy()[:N], g_o.numpy()[:72], N)
    _, F_o, _ = o.energy_forces(cfg, p, R, z, ii, jj)
    np.testing.assert_allclose(Fo.cpu().numpy(), 1.7 * F_o.numpy(), rtol=2e-5, atol=1e-5 if seam_flags == 0 else F_ATOL)
    ii2, jj2 = ii.copy(), jj.copy()
    ii2[5] = jj2[5] = -1
    for flags in (0, 3):                # an asymmetric list is flagged, never a crash, in every arithmetic mode
        eng2 = make(cfg, p, flags=flags)
        eng2.energy_forces(dev(R[None], torch.float32), dev(z[None], torch.int32), dev(ii2[None], torch.int32),
                           dev(jj2[None], torch.int32))
        with pytest.raises(So3kError):
            eng2.check_status()


def test_reference_api_mirror(ethanol):
    # the drop-in surface: So3krates(...), net.init, get_obs_and_force_fn (reference tests/test_model_invariance.py:62-95)
    from mlff_b200.nn import So3krates, get_obs_and_force_fn
    from mlff_b200 import params as P
    net = So3krates(F=36, n_layer=2)
    variables = net.init(0)
    obs_fn = get_obs_and_force_fn(net)
    B = 3
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    out = obs_fn(variables, {'R': R, 'z': z, 'idx_i': nl['idx_i'], 'idx_j': nl['idx_j']})
    assert out['E'].shape == (B, 1) and out['F'].shape == (B, 9, 3)
    flat = P.from_flax_tree(net.cfg, variables)
    cfg = o.OracleConfig(F=36, n_layers=2)
    Eo, Fr = o.energy_forces_batch(cfg, {k: v.astype(np.float64) for k, v in flat.items()}, R, z, nl['idx_i'], nl['idx_j'])
    np.testing.assert_allclose(out['E'], Eo.numpy(), rtol=E_RTOL)
    assert np.abs(out['F'] - Fr.numpy()).max() < F_ATOL
    single = obs_fn(variables, {'R': R[0], 'z': z[0], 'idx_i': nl['idx_i'][0], 'idx_j': nl['idx_j'][0]})
    assert single['E'].shape == (1,) and single['F'].shape == (9, 3)
    np.testing.assert_allclose(single['F'], out['F'][0], atol=1e-7)


# ---- neighbour list ---------------------------------------------------------------------------------------
def gpu_nl(eng, pos, cell, pbc, cutoff, capacity, z=None, mode=0):
    ii, jj, S, cnt = eng.neighbor_list(dev(pos, torch.float64), cutoff, capacity, cell=cell, pbc=pbc,
                                       z=None if z is None else dev(z, torch.int32), mode=mode)
    return ii.cpu().numpy(), jj.cpu().numpy(), S.cpu().numpy(), int(cnt.item())


@pytest.mark.parametrize('pbc', [(1, 1, 1), (1, 1, 0), (0, 0, 0), (1, 0, 1)])
def test_neighbor_list_bit_exact(solid, pbc):
    cfg = o.OracleConfig(F=36, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    oi, oj, oS = o.primitive_neighbor_list(solid['R'], solid['unit_cell'], pbc, 5.0)
    ii, jj, S, cnt = gpu_nl(eng, solid['R'], solid['unit_cell'], pbc, 5.0, len(oi) + 5)
    assert cnt == len(oi)
    assert (ii[:cnt] == oi).all() and (jj[:cnt] == oj).all() and (S[:cnt] == oS).all()
    assert (ii[cnt:] == -1).all() and (S[cnt:] == 0).all()
    assert eng.check_status() == 0


@pytest.mark.parametrize('pbc', [(1, 1, 1), (1, 1, 0)])
def test_minimum_image_list_matches_glp_semantics(solid, pbc):
    """SURVEY 8(a) row a4 (mlff/mdx/atoms.py:488-529, mlff/md/calculator.py:204-238): device mode SO3K_NL_MIC against
    oracle.mic_neighbor_list on a cell smaller than 2 r_cut, on a skewed cell and on a 3 000-atom box where the
    minimum-image and the all-image lists coincide; then through the MD seam."""
    from mlff_b200.capi import NL_MIC
    from common import check_mic_list, jittered_lattice
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p)
    R, cell = solid['R'], solid['unit_cell']
    n_all = len(o.primitive_neighbor_list(R, cell, pbc, 5.0)[0])
    ii, jj, S, cnt = gpu_nl(eng, R, cell, pbc, 5.0, n_all, mode=NL_MIC)
    oi, oj, odv = check_mic_list(o, R, cell, pbc, 5.0, ii, jj, S, cnt)
    assert cnt < n_all and eng.check_status() == 0
    rng = np.random.default_rng(1)
    cell3 = np.array([[7.0, 0, 0], [2.0, 6.5, 0], [1.0, -1.5, 8.0]])
    p3 = rng.uniform(-1, 2, (150, 3)) @ cell3
    ii, jj, S, cnt = gpu_nl(eng, p3, cell3, (1, 1, 1), 3.0, 20000, mode=NL_MIC)
    check_mic_list(o, p3, cell3, (1, 1, 1), 3.0, ii, jj, S, cnt)
    if pbc == (1, 1, 1):
        box = 31.072
        pw = jittered_lattice(3000, box, 2)
        ai, aj, aS = gpu_nl(eng, pw, np.eye(3) * box, pbc, 5.0, 200000)[:3]
        mi, mj, mS, mc = gpu_nl(eng, pw, np.eye(3) * box, pbc, 5.0, 200000, mode=NL_MIC)
        assert (ai[:mc] == mi[:mc]).all() and (aj[:mc] == mj[:mc]).all() and (aS[:mc] == mS[:mc]).all() and ai[mc] == -1
        ea, dE, Fo = eng.potential(dev(odv, torch.float32), dev(solid['z'], torch.int32), dev(oi, torch.int32),
                                   dev(oj, torch.int32))
        ea_o, dE_o = o.potential_displacements(cfg, p, odv, solid['z'], oi, oj, np.ones(len(oi), bool))
        assert eng.check_status() == 0
        np.testing.assert_allclose(ea.cpu().numpy(), ea_o.numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(dE.cpu().numpy(), dE_o.numpy(), rtol=5e-5, atol=5e-5)


def test_neighbor_list_variants(solid, ethanol):
    from mlff_b200.capi import NL_DENSE, STATUS_OVERFLOW
    cfg = o.OracleConfig(F=36, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    ii, jj, S, cnt = gpu_nl(eng, solid['R'], solid['unit_cell'], (1, 1, 1), 5.0, 100)
    assert cnt == 5248 and (eng.check_status() & STATUS_OVERFLOW)
    rng = np.random.default_rng(0)
    cell3 = np.array([[7.0, 0, 0], [2.0, 6.5, 0], [1.0, -1.5, 8.0]])
    p3 = rng.uniform(-1, 2, (200, 3)) @ cell3
    oi, oj, oS = o.primitive_neighbor_list(p3, cell3, (1, 1, 1), 4.0)
    ii, jj, S, cnt = gpu_nl(eng, p3, cell3, (1, 1, 1), 4.0, len(oi) + 3)
    assert cnt == len(oi) and (ii[:cnt] == oi).all() and (jj[:cnt] == oj).all() and (S[:cnt] == oS).all()
    R, z = ethanol['R'][0], ethanol['z']
    nl = o.get_indices(R[None], z[None], 5.0)
    ii, jj, S, cnt = gpu_nl(eng, R, None, (0, 0, 0), 5.0, 80, z=z, mode=NL_DENSE)
    assert cnt == 72 and (ii[:72] == nl['idx_i'][0]).all() and (jj[:72] == nl['idx_j'][0]).all()


@pytest.mark.parametrize('flags', [0, 3])          # fp32 CUDA-core mode ; tensor mode + kept filters (the bench default)
def test_water_box_3000_atoms_vs_oracle(flags):
    # config 3 of BASELINE.json: 3 000 atoms, cubic 31.072 A cell, MIC neighbour list at 5 A, E+F on one GPU.
    # The oracle (float64, brute-force list) still finishes in seconds at this size.
    N, box = 3000, 31.072
    pos = jittered_lattice(N, box, seed=2)
    z = np.where(np.arange(N) % 3 == 0, 8, 1)
    cell = np.eye(3) * box
    cfg = o.OracleConfig()
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=flags)
    oi, oj, oS = o.periodic_neighbor_list_kdtree(pos, box, 5.0)
    ii, jj, S, cnt = gpu_nl(eng, pos, cell, (1, 1, 1), 5.0, int(1.1 * len(oi)))
    assert cnt == len(oi) and (ii[:cnt] == oi).all() and (jj[:cnt] == oj).all() and (S[:cnt] == oS).all()
    E, F, ea, st = run_ef(eng, pos[None], z[None], ii[None], jj[None], cell=cell[None], off=S[None])
    assert st == 0
    Eo, Fo, eao = o.energy_forces(cfg, p, pos, z, oi, oj, cell=cell, cell_offset=oS)
    assert abs(E[0] - Eo.item()) <= E_RTOL * abs(Eo.item())
    assert np.abs(F[0] - Fo.numpy()).max() < F_ATOL
    # size-independent properties: translation invariance, net force ~ 0 (Newton's third law)
    assert np.abs(F[0].sum(0)).max() < 1e-3


@pytest.mark.parametrize('flags', [0, 1, 3])
def test_full_size_properties_100k_atoms(flags):
    # config 4 of BASELINE.json at full size (100 000 atoms, 100 A box): too big for the oracle, so check
    # properties: list symmetric + sorted + strict cutoff, E invariant and F unchanged under a rigid translation
    # through the periodic boundary, net force ~ 0, finite outputs; a 3 000-atom sub-problem is pinned above.
    N, box = 100_000, 100.0
    pos = jittered_lattice(N, box, seed=3)
    z = water_like_numbers(N, 3)
    cell = np.eye(3) * box
    cfg = o.OracleConfig()
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=flags)
    cap = 64 * N
    ii, jj, S, cnt = gpu_nl(eng, pos, cell, (1, 1, 1), 5.0, cap)
    assert 0 < cnt <= cap and eng.check_status() == 0
    ii, jj, S = ii[:cnt], jj[:cnt], S[:cnt]
    oi, oj, oS = o.periodic_neighbor_list_kdtree(pos, box, 5.0)          # bit-exact (i,j,S) set at full size
    assert cnt == len(oi) and (ii == oi).all() and (jj == oj).all() and (S == oS).all()
    assert (np.diff(ii) >= 0).all()
    d = np.linalg.norm(pos[jj] - pos[ii] + S @ cell, axis=-1)
    assert (d < 5.0).all() and (d > 0).all()
    key = lambda a, b, s: (a.astype(np.int64) * N + b) * 27 + (s[:, 0] + 1) * 9 + (s[:, 1] + 1) * 3 + (s[:, 2] + 1)
    assert np.array_equal(np.sort(key(ii, jj, S)), np.sort(key(jj, ii, -S)))
    E, F, ea, st = run_ef(eng, pos[None], z[None], ii[None], jj[None], cell=cell[None], off=S[None])
    assert st == 0 and np.isfinite(E).all() and np.isfinite(F).all()
    assert np.abs(F[0].sum(0)).max() < 2e-2
    pos2 = pos + np.array([3.7, -41.3, 12.9])
    ii2, jj2, S2, cnt2 = gpu_nl(eng, pos2, cell, (1, 1, 1), 5.0, cap)
    assert cnt2 == cnt
    E2, F2, _, _ = run_ef(eng, pos2[None], z[None], ii2[None, :cnt2], jj2[None, :cnt2], cell=cell[None], off=S2[None, :cnt2])
    assert abs(E2[0] - E[0]) <= 2e-5 * abs(E[0])
    assert np.abs(F2[0] - F[0]).max() < 2e-4


def test_calculator_end_to_end(solid):
    from mlff_b200.calculator import mlffCalculator
    from mlff_b200.config import So3kratesConfig
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    calc = mlffCalculator(_cfgs(cfg), {k: np.asarray(v, np.float32) for k, v in p.items()})
    calc.capacity = 1000                                     # force the overflow / re-allocation branch
    res = calc.calculate(solid['R'], solid['z'], cell=solid['unit_cell'], pbc=(True, True, True))
    assert calc.n_edges == 5248 and calc.capacity >= 5248
    assert abs(res['energy'] - float(g['E'])) <= E_RTOL * abs(float(g['E']))
    assert np.abs(res['forces'] - g['F']).max() < F_ATOL


def test_calculator_with_stress(solid):
    """mlffCalculator(calculate_stress=True), mlff/md/calculator.py:87-99: energy, forces and dE/d(strain) in one call."""
    from mlff_b200.calculator import mlffCalculator
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    calc = mlffCalculator(_cfgs(cfg), {k: np.asarray(v, np.float32) for k, v in p.items()}, calculate_stress=True)
    res = calc.calculate(solid['R'], solid['z'], cell=solid['unit_cell'], pbc=(True, True, True))
    _, _, So = o.energy_forces_stress(cfg, p, solid['R'], solid['z'], g['idx_i'], g['idx_j'], cell=solid['unit_cell'],
                                      cell_offset=g['shifts'])
    assert abs(res['energy'] - float(g['E'])) <= E_RTOL * abs(float(g['E']))
    assert np.abs(res['forces'] - g['F']).max() < F_ATOL
    assert res['stress'].shape == (3, 3)
    assert np.abs(calc.get_stress() - So.numpy()).max() < 2e-4 * max(np.abs(So.numpy()).max(), 1.0)
    res = calc.calculate(solid['R'], solid['z'])                      # no cell: no stress entry
    assert 'stress' not in res


# ---- tcgen05 operand-layout conventions ------------------------------------------------------------------------
@pytest.mark.parametrize('N,K,variant', [(144, 32, 0), (144, 176, 0), (176, 144, 2), (256, 64, 0), (48, 144, 2),
                                         (128, 96, 5), (64, 128, 5)])
def test_tcgen05_tile_gemm_layouts(N, K, variant):
    """One 128 x N x K tile on the tensor cores with the canonical no-swizzle operand images the edge kernels
    build (K-major, and the same image consumed MN-major), single pass and 3-pass split-bf16.  Variant 5: the A operand
    comes from tensor memory in the layout the filter kernel's first epilogue writes in place (tcgen05.st, packed
    bf16 hi | lo per 16-column k-step)."""
    from mlff_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.normal(size=(128, K)).astype(np.float32)
    kmajor = variant in (0, 5)
    Bm = rng.normal(size=(N, K) if kmajor else (K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ (Bm.astype(np.float64).T if kmajor else Bm.astype(np.float64))
    def bf(a):      # the kernels' hi part: round-to-nearest-even to bf16 (so3k_tc.cuh split2: cvt.rn.bf16x2.f32)
        b = np.ascontiguousarray(a, np.float32).view(np.uint32)
        b = b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
        return (b & np.uint32(0xFFFF0000)).view(np.float32).astype(np.float64)
    ref1 = bf(A) @ (bf(Bm).T if kmajor else bf(Bm))
    Ad, Bd = dev(A, torch.float32), dev(Bm, torch.float32)
    for passes, target, tol in ((1, ref1, 2e-5), (3, ref, 3e-5)):
        D = torch.zeros(128, N, dtype=torch.float32, device='cuda:0')
        lib.tc_selftest(N, K, variant, passes, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        scale = np.abs(A).astype(np.float64) @ (np.abs(Bm).astype(np.float64).T if kmajor else np.abs(Bm))
        err = np.abs(D.cpu().numpy() - target) / scale
        assert err.max() < tol, (passes, float(err.max()))


# ---- tensor-core (tcgen05, split-bf16) filter GEMMs ------------------------------------------------------------
@pytest.mark.parametrize('keep', [0, 2])
@pytest.mark.parametrize('F,L,degrees,H', [(132, 3, (1, 2, 3), 4), (36, 2, (1, 2, 3), 4), (32, 2, (1, 2, 2, 3), 4)])
def test_tensor_mode_matches_oracle(ethanol, solid, F, L, degrees, H, keep):
    from mlff_b200.capi import FLAG_TENSOR as _FT
    FLAG_TENSOR = _FT | keep          # keep = FLAG_KEEP_FILTERS: the backward streams the forward's filters
    cfg = o.OracleConfig(F=F, n_layers=L, degrees=degrees, n_heads=H)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p, flags=FLAG_TENSOR)
    B = 32
    R = ethanol['R'][:B]
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    E, Fo, ea, st = run_ef(eng, R, z, nl['idx_i'], nl['idx_j'])
    assert st == 0
    Eo, Fr = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(Fo - Fr.numpy()).max() < F_ATOL
    assert np.abs(Fo - Fr.numpy()).max() < 1e-4 * np.abs(Fr.numpy()).max()
    if F == 132:      # periodic solid: many tiles, partial last tile, several images per pair
        g = golden('oracle_solid_pbc.npz')
        eng2 = make(cfg, o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale'])), flags=FLAG_TENSOR)
        E, F_, ea, st = run_ef(eng2, solid['R'][None], solid['z'][None], g['idx_i'][None], g['idx_j'][None],
                               cell=solid['unit_cell'][None], off=g['shifts'][None])
        assert st == 0
        assert abs(E[0] - float(g['E'])) <= E_RTOL * abs(float(g['E']))
        assert np.abs(F_[0] - g['F']).max() < F_ATOL


# ---- domain decomposition (virtual ranks on one GPU; the multi-process NCCL exchange is exercised by bench.py --gpus N
# and, on the CPU, by tests/test_dd.py over gloo) ----------------------------------------------------------------------
@pytest.mark.parametrize('flags', [0, 1, 3])
def test_domain_decomposition_virtual_ranks(solid, flags):
    from mlff_b200 import dist as dd
    from mlff_b200 import _lib
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    device = torch.device('cuda:0')
    pos, z = dev(solid['R'], torch.float64), dev(solid['z'], torch.int64)
    ii, jj, S = (dev(g[k], torch.int64) for k in ('idx_i', 'idx_j', 'shifts'))
    world = 3
    owner = dd.assign_owners_slab(pos, solid['unit_cell'], world, axis=2)
    engines, ranks = [], []
    for r in range(world):
        eng = make(cfg, p, flags=flags)
        engines.append(eng)
        rk = dd.DDRank(eng.lib, eng.h, cfg.n_layers, device)
        rk.begin(dd.build_plan(ii, jj, S, owner, r, world), pos, z, solid['unit_cell'])
        ranks.append(rk)
    dd.run_energy_forces(ranks, dd.exchange_inprocess)
    torch.cuda.synchronize()
    E = sum(float(rk.e_atom[:rk.plan.n_own].double().sum()) for rk in ranks)
    F = np.zeros((110, 3), np.float32)
    for rk in ranks:
        assert int(rk.status.item()) == 0
        F[rk.plan.local_atoms[:rk.plan.n_own].cpu().numpy()] = rk.forces[:rk.plan.n_own].cpu().numpy()
    assert abs(E - float(g['E'])) <= E_RTOL * abs(float(g['E']))
    assert np.abs(F - g['F']).max() < F_ATOL


# ---- batched neighbour lists of a training set in one launch (mlff_b200/indexing.py) ----------------------------------
def test_batched_get_indices_matches_reference_semantics(ethanol):
    from mlff_b200.indexing import get_indices
    cfg = o.OracleConfig(F=36, n_layers=1)
    eng = make(cfg, o.init_params(cfg, 0))
    B = 32
    R = ethanol['R'][:B].astype(np.float64)
    z = np.tile(ethanol['z'][None], (B, 1))
    Rp = np.concatenate([R, np.zeros((B, 2, 3))], 1)            # two padding atoms per frame (z = 0, at the origin)
    zp = np.concatenate([z, np.zeros((B, 2), z.dtype)], 1)
    for r_cut in (5.0, 2.0):                                     # 2 A: ragged lists, frames of different length
        Rr = Rp.copy()
        Rr[3] *= 1.3                                             # stretch one frame so that its list is shorter
        ref = o.get_indices(Rr, zp, r_cut)
        out = get_indices(eng, Rr, zp, r_cut)
        assert out['idx_i'].shape == ref['idx_i'].shape
        assert (out['idx_i'].cpu().numpy() == ref['idx_i']).all() and (out['idx_j'].cpu().numpy() == ref['idx_j']).all()


# ---- stress = dE/d(strain) (observable_function.py:152-186) ---------------------------------------------------------------
@pytest.mark.parametrize('flags', [0, 3])
def test_stress_matches_oracle_strain_derivative(solid, ethanol, flags):
    cfg = o.OracleConfig()
    g = golden('oracle_solid_pbc.npz')
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    eng = make(cfg, p, flags=flags)
    R, z, cell = solid['R'], solid['z'], solid['unit_cell']
    ii, jj, S = g['idx_i'], g['idx_j'], g['shifts']
    E, F, _, st = eng.energy_forces(dev(R[None], torch.float32), dev(z[None], torch.int32), dev(ii[None], torch.int32),
                                    dev(jj[None], torch.int32), dev(cell[None], torch.float32), dev(S[None], torch.int32),
                                    want_stress=True)
    assert eng.check_status() == 0
    Eo, Fo, So = o.energy_forces_stress(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    st = st[0].cpu().numpy()
    assert np.abs(st - st.T).max() < 1e-6 * np.abs(st).max()
    assert np.abs(st - So.numpy()).max() < 2e-4 * max(np.abs(So.numpy()).max(), 1.0)
    # non-periodic batch: every structure gets its own tensor
    B = 3
    Rb = ethanol['R'][:B]
    zb = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(Rb, zb, 5.0)
    cfg2 = o.OracleConfig(F=36, n_layers=2)
    p2 = o.init_params(cfg2, 0, embed_scale=4.0)
    eng2 = make(cfg2, p2, flags=flags)
    _, _, _, stb = eng2.energy_forces(dev(Rb, torch.float32), dev(zb, torch.int32), dev(nl['idx_i'], torch.int32),
                                      dev(nl['idx_j'], torch.int32), want_stress=True)
    for b in range(B):
        _, _, So = o.energy_forces_stress(cfg2, p2, Rb[b], zb[b], nl['idx_i'][b], nl['idx_j'][b])
        assert np.abs(stb[b].cpu().numpy() - So.numpy()).max() < 2e-4 * max(np.abs(So.numpy()).max(), 1.0)


# ---- Energy(zbl_repulsion=True): Ziegler-Biersack-Littmark pair repulsion (observable.py:110-183) ---------------------------
@pytest.mark.parametrize('flags', [0, 3])
def test_zbl_repulsion_matches_oracle(ethanol, flags):
    cfg = o.OracleConfig(F=36, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.37)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    eng = make(cfg, p, flags=flags)
    B = 4
    R = ethanol['R'][:B]                                   # C-H / O-H bonds (~1.0-1.1 A) are inside the 1.5 A switch
    z = np.tile(ethanol['z'][None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    E, F, ea, st = run_ef(eng, R, z, nl['idx_i'], nl['idx_j'])
    assert st == 0
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    cfg0 = o.OracleConfig(F=36, n_layers=2)
    E0, F0 = o.energy_forces_batch(cfg0, p, R, z, nl['idx_i'], nl['idx_j'])
    assert np.abs(Fo.numpy() - F0.numpy()).max() > 1.0     # the repulsion matters in this test
    assert (np.abs(E - Eo.numpy().reshape(-1)) <= E_RTOL * np.abs(Eo.numpy().reshape(-1))).all()
    assert np.abs(F - Fo.numpy()).max() < F_ATOL
    # MD seam ('per_atom' convention: the shift is subtracted from every atom)
    ii, jj = nl['idx_i'][0], nl['idx_j'][0]
    edges = R[0][jj] - R[0][ii]
    mask = np.ones(len(ii), bool)
    ea_g, dE_g, F_g = eng.potential(dev(edges, torch.float32), dev(z[0], torch.int32), dev(ii, torch.int32),
                                    dev(jj, torch.int32), dev(mask, torch.uint8), energy_scale=1.3)
    ea_o, dE_o = o.potential_displacements(cfg, p, edges, z[0], ii, jj, mask, energy_scale=1.3)
    check_md_seam_outputs(flags, ii, jj, ea_g, dE_g, F_g, ea_o.numpy(), dE_o.numpy(), len(z[0]))


@pytest.mark.parametrize('cuda_graph', [False, True], ids=['eager', 'graph'])
def test_dd_session_reuses_its_plan_within_the_skin(cuda_graph):
    """DDSession (mlff_b200/dist.py): the per-rank plan is kept while no atom has moved more than skin / 2 (the skin rule
    of the reference's MD lists, mlff/mdx/atoms.py:365-383); the list is built with r_cut + skin and the extra pairs
    contribute exactly zero, so energies and forces equal a fresh evaluation at the moved positions.  cuda_graph: the
    sweep of a kept plan is captured once and replayed (re-captured after the rebuild)."""
    from mlff_b200 import dist as dd
    from mlff_b200.capi import FLAG_TENSOR, FLAG_KEEP_FILTERS
    from common import jittered_lattice, water_like_numbers
    n, box = 800, 20.0
    pos = jittered_lattice(n, box, 7)
    z = water_like_numbers(n, 7)
    cell = np.eye(3) * box
    cfg = o.OracleConfig(F=132, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=FLAG_TENSOR | FLAG_KEEP_FILTERS)
    sess = dd.DDSession(eng, cfg.n_layers, cell, 1, 0, cfg.r_cut, skin=0.3, cuda_graph=cuda_graph)
    rng = np.random.default_rng(0)
    z_d = dev(z, torch.int32)
    for step, amp in enumerate((0.0, 0.05, 0.05, 0.4)):           # the last move exceeds skin / 2: the plan is rebuilt
        pos = pos + rng.uniform(-amp, amp, pos.shape) / np.sqrt(3.0)
        e, f, ids = sess.energy_forces(dev(pos, torch.float64), z_d)
        ii, jj, S = o.periodic_neighbor_list_kdtree(pos, box, cfg.r_cut)
        Eo, Fo, _ = o.energy_forces(cfg, p, pos, z, ii, jj, cell=cell, cell_offset=S)
        assert abs(float(e.cpu()) - float(Eo)) <= E_RTOL * abs(float(Eo)), step
        F = np.zeros((n, 3), np.float32)
        F[ids.cpu().numpy()] = f.cpu().numpy()
        assert np.abs(F - Fo.numpy()).max() < F_ATOL, step
        assert sess.n_builds == (1 if step < 3 else 2), (step, sess.n_builds)
    assert eng.check_status() == 0


def test_md_through_the_domain_decomposed_calculator():
    """md.CalculatorDD (one rank here: the whole sweep -- plan with a skin, owned / ghost order, stepwise stages -- without the
    collective): a velocity-Verlet trajectory equals the one driven by the single-domain CalculatorX on the same engine."""
    from mlff_b200 import dist as dd
    from mlff_b200 import md
    from mlff_b200.capi import FLAG_TENSOR, FLAG_KEEP_FILTERS
    from common import jittered_lattice, water_like_numbers
    n, box = 600, 18.2
    pos = jittered_lattice(n, box, 9)
    z = water_like_numbers(n, 9)
    cell = np.eye(3) * box
    cfg = o.OracleConfig(F=132, n_layers=2)
    p = o.init_params(cfg, 0, embed_scale=4.0)
    eng = make(cfg, p, flags=FLAG_TENSOR | FLAG_KEEP_FILTERS)
    masses = torch.as_tensor(np.where(z == 8, 15.999, np.where(z == 6, 12.011, 1.008)), device='cuda:0')
    mom = md.maxwell_boltzmann_momenta(masses, 300.0, seed=2)
    z_d = dev(z, torch.int32)
    traj = {}
    for kind in ('single', 'dd'):
        atoms = md.AtomsX(positions=dev(pos, torch.float64), momenta=mom.clone(), masses=masses, numbers=z_d, cell=cell,
                          pbc=(True, True, True))
        if kind == 'single':
            atoms = atoms.init_spatial_partitioning(eng, cfg.r_cut, 0.4)
            calc = md.CalculatorX(eng)
        else:
            sess = dd.DDSession(eng, cfg.n_layers, cell, 1, 0, cfg.r_cut, skin=0.4, cuda_graph=True)
            atoms = atoms.init_domain_decomposition(sess)
            calc = md.CalculatorDD(sess)
        integ = md.VelocityVerletX.create(0.5 * md.FS, calc)
        atoms, st = md.simulate(integ, atoms, 12)
        traj[kind] = (atoms.positions.cpu().numpy(), st['potential_energy'])
    assert eng.check_status() == 0
    assert np.abs(traj['dd'][0] - traj['single'][0]).max() < 1e-5
    assert np.abs(traj['dd'][1] - traj['single'][1]).max() <= 2e-5 * np.abs(traj['single'][1]).max()

