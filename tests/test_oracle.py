"""The oracle against everything the reference pins for this path and against its
own committed golden outputs."""
import math

import numpy as np
import torch
from scipy.spatial.transform import Rotation

import so3k_oracle as o
from conftest import golden


def test_l0_coefficients_match_reference_cgmatrix():
    # mlff/sph_ops/contract.py:162-177 reads diag CG(l,l->0) from cgmatrix.npz; golden/l0_cg_diag.npy holds it
    diag = golden('l0_cg_diag.npy')
    for degrees in ([1, 2, 3], [1, 2, 3, 4], [1, 2, 2, 3], [2, 1]):
        cg, seg = o.l0_coefficients(degrees)
        ref = []
        for d, r in zip(*np.unique(np.array(degrees), return_counts=True)):
            ref += [np.tile(diag[d * d:(d + 1) * (d + 1)], r)]
        np.testing.assert_allclose(cg, np.concatenate(ref), rtol=0, atol=1e-15)
        assert list(seg) == [n for n in range(len(degrees)) for _ in range(2 * degrees[n] + 1)]


def test_ethanol_has_72_edges_per_frame(ethanol):
    # fixture property recorded in SURVEY.md section 4 (r_cut = 5 A)
    z = np.tile(ethanol['z'][None], (32, 1))
    nl = o.get_indices(ethanol['R'], z, 5.0)
    assert nl['idx_i'].shape == (32, 72)
    assert (nl['idx_i'] >= 0).all()


def test_dense_neighbor_list_distance_multiset(solid):
    # reference tests/test_neighbor_list.py:20-49: distances of the list == all pair distances within the cutoff
    R, z = solid['R'][None], solid['z'][None]
    nl = o.get_indices(R, z, 4.0)
    ii, jj = nl['idx_i'][0], nl['idx_j'][0]
    d_list = np.sort(np.linalg.norm(R[0][jj] - R[0][ii], axis=-1))
    D = np.linalg.norm(R[0][:, None] - R[0][None], axis=-1)
    d_all = np.sort(D[(D <= 4.0) & (D > 0)])
    np.testing.assert_allclose(d_list, d_all, rtol=1e-7)


def test_primitive_neighbor_list_properties(solid):
    pos, cell = solid['R'], solid['unit_cell']
    ii, jj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 5.0)
    d = np.linalg.norm(pos[jj] - pos[ii] + S @ cell, axis=-1)
    assert (d < 5.0).all() and (d > 0).all()
    # symmetric: (i,j,S) <-> (j,i,-S)
    fwd = set(zip(ii.tolist(), jj.tolist(), map(tuple, S.tolist())))
    assert all((j, i, (-s[0], -s[1], -s[2])) in fwd for i, j, s in fwd)
    # non-periodic == dense list with strict cutoff
    ii0, jj0, S0 = o.primitive_neighbor_list(pos, cell, [False] * 3, 5.0)
    assert (S0 == 0).all()
    D = np.linalg.norm(pos[:, None] - pos[None], axis=-1)
    assert len(ii0) == int(((D < 5.0) & (D > 0)).sum())


def test_kdtree_list_equals_brute_force():
    from common import jittered_lattice
    pos = jittered_lattice(600, 18.2, seed=5) + np.array([40.0, -3.0, 7.5])      # also outside the box
    a = o.primitive_neighbor_list(pos, np.eye(3) * 18.2, (1, 1, 1), 5.0)
    b = o.periodic_neighbor_list_kdtree(pos, 18.2, 5.0)
    assert all((x == y).all() for x, y in zip(a, b))


def _setup(ethanol, F=36, L=2, degrees=(1, 2, 3), H=4, es=4.0, seed=0):
    cfg = o.OracleConfig(F=F, n_layers=L, degrees=degrees, n_heads=H)
    p = o.init_params(cfg, seed, embed_scale=es)
    R = ethanol['R'][:1]
    z = ethanol['z'][None]
    nl = o.get_indices(R, z, 5.0)
    return cfg, p, R[0], z[0], nl['idx_i'][0], nl['idx_j'][0]


def test_rotation_invariance_equivariance(ethanol):
    # reference tests/test_model_invariance.py:80-130 (F=36, 2 layers ; and repeated degrees [1,2,2,3], F=32)
    for kw in (dict(F=36), dict(F=32, degrees=(1, 2, 2, 3))):
        cfg, p, R, z, ii, jj = _setup(ethanol, **kw)
        E, F, _ = o.energy_forces(cfg, p, R, z, ii, jj, dtype=torch.float32)
        for k in range(3):
            Q = Rotation.random(random_state=k).as_matrix()
            Er, Fr, _ = o.energy_forces(cfg, p, R @ Q.T, z, ii, jj, dtype=torch.float32)
            assert np.isclose(Er.item(), E.item())
            assert not np.isclose(Fr.numpy().reshape(-1), F.numpy().reshape(-1)).all()
            assert np.isclose((F.double() @ torch.tensor(Q.T)).numpy().reshape(-1), Fr.numpy().reshape(-1), atol=1e-5).all()


def test_padding_invariance(ethanol):
    # reference tests/test_model_invariance.py:133-222
    cfg, p, R, z, ii, jj = _setup(ethanol)
    E, F, _ = o.energy_forces(cfg, p, R, z, ii, jj)
    zp = np.concatenate([z, [0, 0, 0]])
    Rp = np.concatenate([R, np.zeros((3, 3))])
    iip = np.concatenate([ii, [-1] * 5])
    jjp = np.concatenate([jj, [-1] * 5])
    Ep, Fp, _ = o.energy_forces(cfg, p, Rp, zp, iip, jjp)
    assert Ep.item() == E.item()
    assert torch.equal(Fp[:9], F)
    assert (Fp[9:] == 0).all()


def test_forces_match_finite_differences(ethanol):
    cfg, p, R, z, ii, jj = _setup(ethanol)
    E, F, _ = o.energy_forces(cfg, p, R, z, ii, jj)
    h = 1e-5
    for (a, c) in ((0, 0), (3, 1), (8, 2)):
        Rp, Rm = R.copy(), R.copy()
        Rp[a, c] += h
        Rm[a, c] -= h
        fd = -(o.energy_forces(cfg, p, Rp, z, ii, jj)[0] - o.energy_forces(cfg, p, Rm, z, ii, jj)[0]).item() / (2 * h)
        assert abs(fd - F[a, c].item()) < 1e-7 * max(1.0, abs(fd))


def test_displacement_convention_equals_positions(ethanol):
    # MD seam (mlff_potential.py:95-109): edges = R_j - R_i gives the same per-atom energies; chaining dE/dedges
    # through R_j - R_i reproduces the forces.
    cfg, p, R, z, ii, jj = _setup(ethanol)
    E, F, ea = o.energy_forces(cfg, p, R, z, ii, jj)
    edges = R[jj] - R[ii]
    mask = np.ones(len(ii), bool)
    ea2, g = o.potential_displacements(cfg, p, edges, z, ii, jj, mask)
    np.testing.assert_allclose(ea2.numpy(), ea.numpy(), rtol=1e-12, atol=1e-14)
    Fch = np.zeros((len(z), 3))
    np.add.at(Fch, ii, g.numpy())
    np.add.at(Fch, jj, -g.numpy())
    np.testing.assert_allclose(Fch, F.numpy(), rtol=1e-10, atol=1e-13)


def test_oracle_reproduces_committed_golden_outputs(ethanol, solid):
    cfg = o.OracleConfig()
    for tag in ('default', 'amplified'):
        g = golden(f'oracle_ethanol_{tag}.npz')
        p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
        z = np.tile(ethanol['z'][None], (4, 1))
        E, F = o.energy_forces_batch(cfg, p, ethanol['R'][:4], z, g['idx_i'], g['idx_j'])
        np.testing.assert_allclose(E.numpy(), g['E'], rtol=1e-12)
        np.testing.assert_allclose(F.numpy(), g['F'], rtol=1e-9, atol=1e-13)


def test_stress_is_the_strain_derivative_of_the_energy():
    # observable_function.py:152-186: stress = dE/d eps for R -> R + R eps_s^T, cell -> cell + cell eps_s^T;
    # check the autograd result against central finite differences in float64 and against the virial form
    # sym(sum_e dE/dr_e (x) r_e) that the CUDA kernel evaluates
    g = golden('oracle_solid_pbc.npz')
    sol = golden('solid_0.npz')
    cfg = o.OracleConfig(F=36, n_layers=2)
    p = o.init_params(cfg, 3, embed_scale=4.0)
    R, z, cell = sol['R'], sol['z'], sol['unit_cell']
    ii, jj, S = g['idx_i'], g['idx_j'], g['shifts']
    E, F, st = o.energy_forces_stress(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    st = st.numpy()
    assert np.abs(st - st.T).max() < 1e-12 * max(np.abs(st).max(), 1.0)
    h = 1e-6
    for a, b in ((0, 0), (0, 1), (2, 1)):
        eps = np.zeros((3, 3))
        eps[a, b] = h
        es = 0.5 * (eps + eps.T)
        Ep, _, _ = o.energy_forces(cfg, p, R + R @ es.T, z, ii, jj, cell=cell + cell @ es.T, cell_offset=S)
        Em, _, _ = o.energy_forces(cfg, p, R - R @ es.T, z, ii, jj, cell=cell - cell @ es.T, cell_offset=S)
        fd = (Ep.item() - Em.item()) / (2 * h)
        assert abs(fd - st[a, b]) < 1e-6 * max(np.abs(st).max(), 1.0)
    # virial form over the directed edges
    r = R[jj] - R[ii] + S @ cell
    e_atom, dE = o.potential_displacements(cfg, p, r, z, ii, jj, np.ones(len(ii), bool))
    vir = dE.numpy().T @ r
    np.testing.assert_allclose(0.5 * (vir + vir.T), st, atol=1e-9 * max(np.abs(st).max(), 1.0))


def test_oracle_matches_reference_source_functions():
    """The stateless pieces of the path against arrays produced by the REFERENCE'S OWN SOURCE FILES (executed by
    tests/golden/make_reference_function_golden.py with numpy standing in for jax.numpy): GeometryEmbed.__call__ in
    the positions, periodic (cell offsets) and displacements conventions with padded pairs, an out-of-cutoff pair and a
    zero vector -- r_ij, d_ij, PhysNet basis, cosine cutoff, unit vectors, real spherical harmonics l = 0..4 -- and
    make_l0_contraction_fn for three degree lists.  This pins the oracle's formulas and masks to the reference itself."""
    g = golden('reference_functions.npz')
    cfg = o.OracleConfig(degrees=(0, 1, 2, 3, 4), n_rbf=32, r_cut=5.0)
    t = lambda a: torch.as_tensor(np.asarray(a))
    for tag in ('molecule', 'periodic', 'displacements'):
        ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith(tag + '/in/')}
        kw = {}
        if tag == 'displacements':
            kw['displacements'] = t(ins['displacements'])
        else:
            kw['R'] = t(ins['R'])
        if tag == 'periodic':
            kw['cell'], kw['cell_offset'] = t(ins['unit_cell']), t(ins['cell_offset'])
        geo = o.geometry_embed(cfg, t(ins['idx_i']).long(), t(ins['idx_j']).long(), t(ins['pair_mask']), torch.float64, **kw)
        n_checked = 0
        for k in ('r_ij', 'd_ij', 'rbf_ij', 'phi_r_cut', 'unit_r_ij', 'sph_ij'):
            ref = g[f'{tag}/{k}']
            assert np.isfinite(ref).all()
            np.testing.assert_allclose(geo[k].numpy(), ref, rtol=1e-12, atol=1e-13, err_msg=f'{tag}/{k}')
            n_checked += ref.size
        assert n_checked > 1000
        assert (g[f'{tag}/chi'] == 0).all()                     # sphc_normalization=None: chi0 = 0 (embed.py:196-197)
    # the fixtures do exercise the special cases
    assert (g['molecule/phi_r_cut'] == 0).sum() >= 5            # the 6.5 A pair (both directions) and the padding slots
    assert (g['displacements/unit_r_ij'][5] == 0).all() and g['displacements/d_ij'][5] == 0
    for tag in ('123', '1223', '01234'):
        degs = tuple(int(x) for x in g[f'l0/{tag}/degrees'])
        out = o.l0_contraction(t(g[f'l0/{tag}/chi']), degs)
        np.testing.assert_allclose(out.numpy(), g[f'l0/{tag}/out'], rtol=1e-12, atol=1e-14)


def test_oracle_matches_reference_model_code():
    """Energies and finite-difference forces produced by the REFERENCE'S OWN MODEL CODE (init_so3krates ->
    StackNet.__call__ -> So3kratesLayer / Energy, executed unmodified in float64 on the jax / flax stand-ins of
    tests/golden/ref_shim.py by tests/golden/make_reference_model_golden.py) against the oracle restatement on the same
    parameters and inputs: default model on ethanol with padding, a repeated-degree list, every optional layer
    branch, the ZBL repulsion, and the periodic solid with cell offsets."""
    g = golden('reference_model.npz')
    cases = {'ethanol_default': (o.OracleConfig(), dict(scale_shift=True)),
             'ethanol_1223': (o.OracleConfig(F=32, n_layers=2, degrees=(1, 2, 2, 3), n_heads=2), {}),
             'ethanol_all_branches': (o.OracleConfig(F=36, n_layers=2, layer_normalization=True, final_layer=True,
                                                     residual_mlp_1=True, residual_mlp_2=True), {}),
             'ethanol_zbl': (o.OracleConfig(F=36, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.37),
                             dict(scale_shift=True)),
             'solid_periodic': (o.OracleConfig(F=36, n_layers=2), {})}
    for tag, (cfg, kw) in cases.items():
        p = o.init_params(cfg, int(g[f'{tag}/seed']), embed_scale=4.0, **kw)
        ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith(tag + '/in/')}
        E, F, _ = o.energy_forces(cfg, p, ins['R'], ins['z'], ins['idx_i'], ins['idx_j'], cell=ins.get('unit_cell'),
                                  cell_offset=ins.get('cell_offset'))
        assert abs(float(E) - float(g[f'{tag}/E'])) <= 1e-11 * max(1.0, abs(float(g[f'{tag}/E']))), tag
        F_fd = g[f'{tag}/F_fd']                                # central differences of the reference energy, h = 1e-4
        np.testing.assert_allclose(F.numpy()[g[f'{tag}/fd_atoms']], F_fd, rtol=1e-6, atol=2e-7, err_msg=tag)
        assert np.abs(F_fd).max() > 0.05


def test_oracle_md_seam_matches_reference_model_code():
    """MLFFPotential.potential_fn (mdx/potential/mlff_potential.py:58-109) replayed on the reference's own model code
    ('displacements' input and 'per_atom' output conventions, graph mask -> idx -1, energy scale, ZBL shift per atom):
    per-atom energies and finite-difference dE/d(edges) against the oracle's potential_displacements."""
    g = golden('reference_model.npz')
    zk = dict(zbl_repulsion=True, zbl_repulsion_shift=0.37)
    for tag, cfg in (('md_seam_default', o.OracleConfig(F=36, n_layers=2)),
                     ('md_seam_zbl', o.OracleConfig(F=36, n_layers=2, **zk))):
        p = o.init_params(cfg, int(g[f'{tag}/seed']), embed_scale=4.0, scale_shift=True)
        ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith(tag + '/in/')}
        assert not ins['mask'].all()
        ea, dE = o.potential_displacements(cfg, p, ins['edges'], ins['z'], ins['centers'], ins['others'], ins['mask'],
                                           energy_scale=float(g[f'{tag}/scale']))
        np.testing.assert_allclose(ea.numpy(), g[f'{tag}/e_atom'], rtol=1e-11, atol=1e-12, err_msg=tag)
        fd = g[f'{tag}/dE_dedges_fd']
        np.testing.assert_allclose(dE.numpy()[g[f'{tag}/fd_edges']], fd, rtol=1e-6, atol=2e-7, err_msg=tag)
        assert (fd[1] == 0).all() and np.abs(fd).max() > 0.05          # edge 3 is masked: no dependence


def test_get_indices_matches_reference_source():
    """mlff/indexing/indices.py:15-84 executed from the reference's source (numpy stand-in for jax.numpy) on a padded
    batch: the oracle's get_indices returns the identical arrays (same row-major order, same -1 padding), including a
    pair at exactly r_cut, an isolated atom and padding atoms."""
    g = golden('reference_functions.npz')
    R, z = g['get_indices/R'], g['get_indices/z']
    for rc in (5.0, 2.0):
        nl = o.get_indices(R, z, rc)
        assert np.array_equal(nl['idx_i'], g[f'get_indices/{rc}/idx_i'])
        assert np.array_equal(nl['idx_j'], g[f'get_indices/{rc}/idx_j'])
    assert (g['get_indices/5.0/idx_i'] >= 9).sum() == 0          # padding atoms (z = 0) never appear
    i3, j3 = g['get_indices/5.0/idx_i'][3], g['get_indices/5.0/idx_j'][3]
    assert ((i3 == 0) & (j3 == 1)).any()                         # d == r_cut exactly is a neighbour (<=)


def test_oracle_stress_matches_reference_stress_function():
    """get_energy_force_stress_fn (mlff/nn/stacknet/observable_function.py:152-186) executed from the reference's source
    (jax.value_and_grad replaced by float64 central differences) on a periodic system: E, forces and the strain
    derivative (the reference's stress convention, no volume factor) against the oracle."""
    g = golden('reference_model.npz')
    cfg = o.OracleConfig(F=32, n_layers=2, degrees=(1, 2), n_heads=2)
    p = o.init_params(cfg, 7, embed_scale=4.0)
    ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith('stress/in/')}
    E, F, stress = o.energy_forces_stress(cfg, p, ins['R'], ins['z'], ins['idx_i'], ins['idx_j'], cell=ins['unit_cell'],
                                          cell_offset=ins['cell_offset'])
    assert abs(float(E) - float(g['stress/E'].reshape(()))) < 1e-11 * abs(float(E))
    np.testing.assert_allclose(F.numpy(), g['stress/F'], rtol=1e-6, atol=5e-7)
    np.testing.assert_allclose(stress.numpy(), g['stress/stress'], rtol=1e-6, atol=5e-7)
    assert np.abs(g['stress/stress']).max() > 1.0
