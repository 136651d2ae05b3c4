"""MD inner loop (mlff_b200/md.py): the integrator / skin-list logic against a plain numpy velocity Verlet
driven by the oracle.  CPU tests use an oracle-backed engine stand-in; GPU tests use the CUDA engine."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle'))
import so3k_oracle as o
from conftest import golden
from mlff_b200 import md

MASS = {1: 1.008, 6: 12.011, 8: 15.999}


class OracleEngine:
    """neighbor_list / energy_forces with the engine's signatures, computed by the oracle in float64 (test stand-in)."""

    def __init__(self, cfg, params):
        self.cfg, self.params = cfg, params

    def neighbor_list(self, positions, cutoff, capacity, cell=None, pbc=(False, False, False), **kw):
        pos = positions.detach().cpu().numpy()
        c = np.eye(3) * 1e3 if cell is None else cell
        i, j, S = o.primitive_neighbor_list(pos, c, pbc, cutoff)
        n = len(i)
        ii = torch.full((capacity,), -1, dtype=torch.int32)
        jj = torch.full((capacity,), -1, dtype=torch.int32)
        SS = torch.zeros((capacity, 3), dtype=torch.int32)
        m = min(n, capacity)
        ii[:m], jj[:m], SS[:m] = torch.as_tensor(i[:m]), torch.as_tensor(j[:m]), torch.as_tensor(S[:m])
        return ii, jj, SS, torch.tensor([n])

    def energy_forces(self, R, z, idx_i, idx_j, cell=None, cell_offset=None, want_stress=False):
        kw = dict(cell=None if cell is None else cell[0].double().numpy(),
                  cell_offset=None if cell_offset is None else cell_offset[0].numpy())
        if want_stress:
            E, F, st = o.energy_forces_stress(self.cfg, self.params, R[0].double().numpy(), z[0].numpy(), idx_i[0].numpy(),
                                              idx_j[0].numpy(), **kw)
            return E.reshape(1, 1), F[None], None, st[None]
        E, F, _ = o.energy_forces(self.cfg, self.params, R[0].double().numpy(), z[0].numpy(), idx_i[0].numpy(),
                                  idx_j[0].numpy(), **kw)
        return E.reshape(1, 1), F[None], None


def reference_trajectory(cfg, params, R0, z, p0, masses, dt, steps, r_cut=5.0):
    """numpy velocity Verlet with the exact-cutoff list rebuilt at every step and oracle forces (float64)."""
    def forces(R):
        nl = o.get_indices(R[None], z[None], r_cut)
        _, F, _ = o.energy_forces(cfg, params, R, z, nl['idx_i'][0], nl['idx_j'][0])
        return F.numpy()
    R, p = R0.copy(), p0.copy()
    F = forces(R)
    for _ in range(steps):
        p = p + 0.5 * dt * F
        R = R + dt * p / masses[:, None]
        F = forces(R)
        p = p + 0.5 * dt * F
    return R, p


def ethanol_state(device='cpu'):
    g = golden('ethanol_32.npz')
    R0, z = g['R'][0].astype(np.float64), g['z'].astype(np.int64)
    masses = np.array([MASS[int(a)] for a in z])
    p0 = md.maxwell_boltzmann_momenta(torch.as_tensor(masses), 300.0, seed=1).numpy()
    atoms = md.AtomsX(positions=torch.as_tensor(R0, device=device), momenta=torch.as_tensor(p0, device=device),
                      masses=torch.as_tensor(masses, device=device), numbers=torch.as_tensor(z, dtype=torch.int32, device=device))
    return R0, z, masses, p0, atoms


def test_velocity_verlet_with_skin_list_matches_plain_integrator():
    cfg = o.OracleConfig(F=36, n_layers=2)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    R0, z, masses, p0, atoms = ethanol_state()
    eng = OracleEngine(cfg, params)
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=1.0)
    calc = md.CalculatorX(eng)
    integ = md.VelocityVerletX.create(0.5 * md.FS, calc)
    steps = 6
    atoms, stats = md.simulate(integ, atoms, steps)
    R_ref, p_ref = reference_trajectory(cfg, params, R0, z, p0, masses, 0.5 * md.FS, steps)
    # skin edges (5 A <= d < 6 A) contribute exactly zero, so the trajectories agree up to the float32 rounding of the
    # positions handed to the calculator (the engine's input type)
    np.testing.assert_allclose(atoms.positions.numpy(), R_ref, atol=1e-7)
    np.testing.assert_allclose(atoms.momenta.numpy(), p_ref, atol=1e-6)
    assert calc.n_calls == steps + 1                      # force re-use: one evaluation per step (+ the initial one)
    assert atoms.neighbors.n_builds == 1                  # nothing moved 0.5 A in 3 fs
    assert stats['kinetic_energy'].shape == (steps,) and np.isfinite(stats['potential_energy']).all()


def test_skin_list_rebuilds_after_half_skin_displacement():
    cfg = o.OracleConfig(F=36, n_layers=1)
    eng = OracleEngine(cfg, None)
    _, _, _, _, atoms = ethanol_state()
    atoms = atoms.init_spatial_partitioning(eng, cutoff=3.0, skin=0.4)
    n0 = atoms.neighbors.count
    small = atoms.positions.clone()
    small[0, 0] += 0.19
    a1 = atoms.update_positions(small)
    assert a1.neighbors.n_builds == 1
    big = atoms.positions.clone()
    big[0, 0] += 0.21
    a2 = atoms.update_positions(big)
    assert a2.neighbors.n_builds == 2 and a2.neighbors.count > 0 and n0 > 0
    i, j, _ = o.primitive_neighbor_list(big.numpy(), np.eye(3) * 1e3, (False,) * 3, 3.4)
    assert a2.neighbors.count == len(i)


@pytest.mark.gpu
@pytest.mark.parametrize('flags', [0, 3])
def test_md_on_device_matches_oracle_trajectory(flags):
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    cfg = o.OracleConfig(F=36, n_layers=2)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    eng = So3kEngine(So3kratesConfig(F=36, n_layer=2), 'cuda:0', flags)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
    R0, z, masses, p0, atoms = ethanol_state('cuda:0')
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=1.0)
    calc = md.CalculatorX(eng)
    steps = 10
    atoms, stats = md.simulate(md.VelocityVerletX.create(0.5 * md.FS, calc), atoms, steps)
    assert eng.check_status() == 0
    R_ref, p_ref = reference_trajectory(cfg, params, R0, z, p0, masses, 0.5 * md.FS, steps)
    assert np.abs(atoms.positions.cpu().numpy() - R_ref).max() < 1e-5
    assert np.abs(atoms.momenta.cpu().numpy() - p_ref).max() < 1e-4
    assert calc.n_calls == steps + 1


@pytest.mark.gpu
def test_md_periodic_box_conserves_energy_and_reuses_the_list():
    from common import jittered_lattice
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    N, box = 1000, 21.5
    pos = jittered_lattice(N, box, seed=4)
    z = np.where(np.arange(N) % 3 == 0, 8, 1)
    masses = np.array([MASS[int(a)] for a in z])
    cfg = o.OracleConfig()
    params = o.init_params(cfg, 0, embed_scale=4.0)
    eng = So3kEngine(So3kratesConfig(n_layer=3), 'cuda:0', 3)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
    dev = 'cuda:0'
    m_t = torch.as_tensor(masses, device=dev)
    atoms = md.AtomsX(positions=torch.as_tensor(pos, device=dev), momenta=md.maxwell_boltzmann_momenta(m_t, 300.0, seed=2),
                      masses=m_t, numbers=torch.as_tensor(z, dtype=torch.int32, device=dev), cell=np.eye(3) * box,
                      pbc=(True, True, True))
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=0.5)
    calc = md.CalculatorX(eng)
    steps = 40
    atoms, stats = md.simulate(md.VelocityVerletX.create(0.25 * md.FS, calc), atoms, steps)
    assert eng.check_status() == 0
    etot = stats['kinetic_energy'] + stats['potential_energy']
    # NVE: the total energy fluctuates by a small fraction of the kinetic energy and does not drift
    assert np.abs(etot - etot[0]).max() < 0.02 * stats['kinetic_energy'].mean()
    assert atoms.neighbors.n_builds < steps // 4          # the skin list is re-used
    assert calc.n_calls == steps + 1


@pytest.mark.gpu
def test_md_cuda_graph_replay_equals_eager():
    # small systems are launch-bound: CalculatorX(cuda_graph=True) replays one captured graph per evaluation
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    cfg = o.OracleConfig(F=36, n_layers=2)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    out = []
    for use_graph in (False, True):
        eng = So3kEngine(So3kratesConfig(F=36, n_layer=2), 'cuda:0', 3)
        eng.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
        _, _, _, _, atoms = ethanol_state('cuda:0')
        atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=1.0)
        calc = md.CalculatorX(eng, cuda_graph=use_graph)
        atoms, stats = md.simulate(md.VelocityVerletX.create(0.5 * md.FS, calc), atoms, 12)
        assert eng.check_status() == 0
        out.append((atoms.positions.cpu().numpy(), stats['potential_energy']))
    np.testing.assert_allclose(out[0][0], out[1][0], atol=1e-9)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-6)



def test_nose_hoover_matches_plain_integrator():
    # NoseHooverX.step (mlff/mdx/integrator.py:49-83) against the same update rule written out in numpy with oracle forces
    cfg = o.OracleConfig(F=36, n_layers=2)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    R0, z, masses, p0, atoms = ethanol_state()
    eng = OracleEngine(cfg, params)
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=1.0)
    dt, T, ttime, steps = 0.5 * md.FS, 300.0 * md.KB_EV, 20.0, 5
    integ = md.NoseHooverX.create(dt, T, ttime, md.CalculatorX(eng))
    for _ in range(steps):
        integ, atoms, _ = integ.step(atoms)

    def forces(R):
        nl = o.get_indices(R[None], z[None], 5.0)
        return o.energy_forces(cfg, params, R.astype(np.float32).astype(np.float64), z, nl['idx_i'][0], nl['idx_j'][0])[1].numpy()
    m = masses[:, None]
    n = len(z)
    R, v, zeta = R0.copy(), p0 / m, 0.0
    target, Q = 0.5 * 3.0 * n * T, 3.0 * n * T * (ttime * dt) ** 2
    ekin = lambda vv: 0.5 * float((m * vv * vv).sum())
    for _ in range(steps):
        a = forces(R) / m
        R = R + v * dt + (a - zeta * v) * (0.5 * dt ** 2)
        ke0 = ekin(v)
        vh = v + 0.5 * dt * (a - zeta * v)
        a = forces(R) / m
        zeta = zeta + 0.5 * dt / Q * (ke0 - target)
        zeta = zeta + 0.5 * dt / Q * (ekin(vh) - target)
        v = (vh + 0.5 * dt * a) / (1 + 0.5 * dt * zeta)
    np.testing.assert_allclose(atoms.positions.numpy(), R, atol=1e-8)
    np.testing.assert_allclose(atoms.get_velocities().numpy(), v, atol=1e-8)
    assert abs(float(integ.zeta) - zeta) < 1e-10


def test_langevin_thermalises_free_particles():
    # no forces: the Langevin step must drive the kinetic temperature to the target and keep the centre of mass fixed
    class FreeEngine:
        def neighbor_list(self, positions, cutoff, capacity, cell=None, pbc=(False,) * 3, **kw):
            z1 = torch.full((capacity,), -1, dtype=torch.int32)
            return z1, z1.clone(), torch.zeros((capacity, 3), dtype=torch.int32), torch.tensor([0])

        def energy_forces(self, R, z, ii, jj, cell=None, off=None):
            return torch.zeros(1, 1), torch.zeros_like(R), None
    n = 2000
    masses = torch.full((n,), 12.0, dtype=torch.float64)
    atoms = md.AtomsX(positions=torch.zeros(n, 3, dtype=torch.float64), momenta=torch.zeros(n, 3, dtype=torch.float64),
                      masses=masses, numbers=torch.full((n,), 6, dtype=torch.int32))
    atoms = atoms.init_spatial_partitioning(FreeEngine(), cutoff=1.0, skin=1e9)
    T = 300.0 * md.KB_EV
    integ = md.LangevinX.create(1.0 * md.FS, T, 0.05 / md.FS, md.CalculatorX(FreeEngine()), seed=3)
    for _ in range(400):
        integ, atoms, _ = integ.step(atoms)
    temp = float(2.0 * atoms.get_kinetic_energy() / (3 * n)) / md.KB_EV
    assert abs(temp - 300.0) < 15.0                                      # 6000 degrees of freedom: ~1.8 % standard deviation
    assert float(atoms.momenta.sum(0).abs().max()) < 1e-9 * n


class SpringEngine:
    """Engine stand-in with a known minimum: E = sum over directed edges of k/4 (d - d0)^2 (each bond counted twice)."""

    def __init__(self, d0=1.2, k=5.0):
        self.d0, self.k = d0, k
        self.neighbor_list = OracleEngine(None, None).neighbor_list

    def energy_forces(self, R, z, idx_i, idx_j, cell=None, cell_offset=None):
        R0 = R[0].double().clone().requires_grad_(True)
        keep = idx_i[0] >= 0
        i, j = idx_i[0][keep].long(), idx_j[0][keep].long()
        d = (R0[j] - R0[i]).norm(dim=-1)
        E = (0.25 * self.k * (d - self.d0) ** 2).sum()
        (g,) = torch.autograd.grad(E, R0)
        return E.detach().reshape(1, 1), (-g).detach()[None], None


def _tetrahedron():
    R0 = np.array([[0.0, 0.0, 0.0], [1.5, 0.1, 0.0], [0.6, 1.3, 0.2], [0.5, 0.4, 1.6]])
    n = len(R0)
    return md.AtomsX(positions=torch.as_tensor(R0), momenta=torch.zeros(n, 3, dtype=torch.float64),
                     masses=torch.ones(n, dtype=torch.float64), numbers=torch.ones(n, dtype=torch.int32))


def test_gradient_descent_relaxer_follows_the_reference_rule():
    # mlff/mdx/optimizer.py:29-52: x <- x - lr * dE/dx ; stops when the largest atomic gradient norm < tol
    eng = SpringEngine()
    atoms = _tetrahedron().init_spatial_partitioning(eng, cutoff=5.0, skin=0.5)
    calc = md.CalculatorX(eng)
    gd = md.GradientDescent.create(calc, learning_rate=5e-2)
    F0 = calc(atoms)['forces'].clone()
    a1, grads = gd.step(atoms)
    np.testing.assert_allclose(grads.numpy(), -F0.numpy(), atol=0)
    np.testing.assert_allclose(a1.get_positions().numpy(), (atoms.get_positions() + 5e-2 * F0).numpy(), atol=1e-15)
    relaxed, grads = gd.minimize(atoms, max_steps=2000, tol=1e-4)
    assert float(torch.linalg.norm(grads, dim=-1).max()) < 1e-4
    d = torch.cdist(relaxed.get_positions(), relaxed.get_positions())[np.triu_indices(4, 1)]
    np.testing.assert_allclose(d.numpy(), 1.2, atol=2e-4)          # regular tetrahedron with the spring length
    with pytest.raises(RuntimeError, match='did not converge'):
        gd.minimize(atoms, max_steps=3, tol=1e-8)


def test_lbfgs_relaxer_converges_and_beats_gradient_descent():
    eng = SpringEngine()
    atoms = _tetrahedron().init_spatial_partitioning(eng, cutoff=5.0, skin=0.5)
    calc = md.CalculatorX(eng)
    opt = md.LBFGS.create(atoms, calc)
    relaxed, forces = opt.minimize(atoms, max_steps=60, tol=1e-6)
    assert float(torch.linalg.norm(forces, dim=-1).max()) < 1e-6
    d = torch.cdist(relaxed.get_positions(), relaxed.get_positions())[np.triu_indices(4, 1)]
    np.testing.assert_allclose(d.numpy(), 1.2, atol=1e-5)
    assert calc.n_calls < 150                                     # gradient descent needs > 300 evaluations for 1e-4
    with pytest.raises(RuntimeError, match='did not converge'):
        md.LBFGS.create(atoms, calc).minimize(atoms, max_steps=1, tol=1e-12)


def test_lbfgs_relaxes_ethanol_on_the_model_surface():
    """The SO3krates surface itself (oracle-backed engine): the energy goes down monotonically and the largest force
    falls by more than an order of magnitude; the skin list is rebuilt on the way when atoms move far enough."""
    cfg = o.OracleConfig(F=32, n_layers=2, degrees=(1, 2), n_heads=4)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    eng = OracleEngine(cfg, params)
    _, _, _, _, atoms = ethanol_state()
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=0.5)
    calc = md.CalculatorX(eng)
    f0 = float(torch.linalg.norm(calc(atoms)['forces'], dim=-1).max())
    opt = md.LBFGS.create(atoms, calc)
    energies = [float(calc(atoms)['energy'])]
    for _ in range(25):
        atoms, grads = opt.step(atoms)
        energies.append(float(calc(atoms)['energy']))
    assert all(b <= a + 1e-12 for a, b in zip(energies, energies[1:]))
    assert float(torch.linalg.norm(grads, dim=-1).max()) < 0.1 * f0


def test_integrators_match_reference_source():
    """VelocityVerletX, NoseHooverX and BerendsenX of mlff_b200.md against trajectories recorded from the REFERENCE'S OWN integrator
    code (mlff/mdx/integrator.py:16-83, 237-310, 312-342, executed by tests/golden/make_reference_md_golden.py on numpy
    stand-ins) for the same spring network, initial state and time step: positions, momenta, potential energy and the
    thermostat variable agree step by step for 50 steps."""
    g = golden('reference_md.npz')
    D0, K = float(g['D0']), float(g['K'])

    class Springs:                                  # the generator's analytic calculator, in torch
        def __call__(self, atoms):
            if atoms.cache is None:
                R = atoms.get_positions()
                d = R[None, :, :] - R[:, None, :]
                eye = torch.eye(len(R), dtype=R.dtype)
                r = d.norm(dim=-1) + eye
                f = K * (r - D0) / r * (1 - eye)
                atoms.cache = {'energy': 0.25 * K * ((r - D0) ** 2 * (1 - eye)).sum(), 'forces': (f[:, :, None] * d).sum(1)}
            return atoms.cache

    class NoList:
        def update(self, positions):
            return self

    def start():
        return md.AtomsX(positions=torch.as_tensor(g['R0']), momenta=torch.as_tensor(g['p0']),
                         masses=torch.as_tensor(g['masses']), numbers=torch.ones(len(g['masses']), dtype=torch.int32),
                         neighbors=NoList())

    dt = float(g['dt'])
    runs = {'velocity_verlet': md.VelocityVerletX.create(dt, Springs()),
            'nose_hoover': md.NoseHooverX.create(dt, float(g['nh_temperature']), float(g['nh_ttime']), Springs()),
            'berendsen': md.BerendsenX.create(dt, float(g['be_temperature']), Springs(), float(g['be_time_constant']))}
    for tag, integrator in runs.items():
        atoms = start()
        for k in range(len(g[f'{tag}/E'])):
            integrator, atoms, prop = integrator.step(atoms)
            np.testing.assert_allclose(atoms.get_positions().numpy(), g[f'{tag}/R'][k], rtol=1e-12, atol=1e-13, err_msg=f'{tag} {k}')
            np.testing.assert_allclose(atoms.get_momenta().numpy(), g[f'{tag}/p'][k], rtol=1e-12, atol=1e-13, err_msg=f'{tag} {k}')
            assert abs(float(prop['energy']) - g[f'{tag}/E'][k]) < 1e-12
            if tag == 'nose_hoover':
                assert abs(float(integrator.zeta) - g[f'{tag}/zeta'][k]) < 1e-12
    assert g['nose_hoover/zeta'][-1] > 1.0                        # the thermostat is active in this run


def test_langevin_update_matches_reference_source(monkeypatch):
    """LangevinX (mlff/mdx/integrator.py:86-159) executed from the reference's source with its jax.random draws recorded;
    mlff_b200.md.LangevinX fed the same xi / eta reproduces the trajectory (coefficients c1..c5, centre-of-mass fix,
    RATTLE-style velocity recomputation) to 1e-12."""
    g = golden('reference_md.npz')
    D0, K = float(g['D0']), float(g['K'])
    stream = iter(g['langevin/noise'])
    monkeypatch.setattr(md.torch, 'randn', lambda *shape, **kw: torch.as_tensor(next(stream)))

    class Springs:
        def __call__(self, atoms):
            if atoms.cache is None:
                R = atoms.get_positions()
                d = R[None, :, :] - R[:, None, :]
                eye = torch.eye(len(R), dtype=R.dtype)
                r = d.norm(dim=-1) + eye
                f = K * (r - D0) / r * (1 - eye)
                atoms.cache = {'energy': 0.25 * K * ((r - D0) ** 2 * (1 - eye)).sum(), 'forces': (f[:, :, None] * d).sum(1)}
            return atoms.cache

    class NoList:
        def update(self, positions):
            return self

    atoms = md.AtomsX(positions=torch.as_tensor(g['R0']), momenta=torch.as_tensor(g['p0']), masses=torch.as_tensor(g['masses']),
                      numbers=torch.ones(len(g['masses']), dtype=torch.int32), neighbors=NoList())
    integrator = md.LangevinX.create(float(g['dt']), float(g['lv_temperature']), float(g['lv_friction']), Springs())
    for k in range(len(g['langevin/E'])):
        integrator, atoms, prop = integrator.step(atoms)
        np.testing.assert_allclose(atoms.get_positions().numpy(), g['langevin/R'][k], rtol=1e-12, atol=1e-13, err_msg=str(k))
        np.testing.assert_allclose(atoms.get_momenta().numpy(), g['langevin/p'][k], rtol=1e-12, atol=1e-13, err_msg=str(k))


# ---- momentum utilities (mlff/mdx/utils.py), after the reference's tests/test_mdx_utils.py ---------------------------------
def _utils_atoms(g, p=None):
    return md.AtomsX(positions=torch.as_tensor(g['R']), momenta=torch.as_tensor(g['p0'] if p is None else p),
                     masses=torch.as_tensor(g['masses']), numbers=torch.as_tensor(g['z']))


def test_momentum_utilities_match_reference_source():
    """zero_translation, zero_rotation, scale_momenta and the AtomsX accessors under them against values recorded from the
    REFERENCE'S OWN code (mlff/mdx/utils.py:6-83, mlff/mdx/atoms.py:211-288, 385-441, executed by
    tests/golden/make_reference_mdx_utils_golden.py) on the ethanol frame of the reference's test, 500 K momenta."""
    g = golden('reference_mdx_utils.npz')
    a0 = _utils_atoms(g)
    tol = dict(rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(float(a0.get_temperature()), g['temperature'], **tol)
    np.testing.assert_allclose(a0.get_center_of_mass().numpy(), g['center_of_mass'], **tol)
    np.testing.assert_allclose(a0.get_center_of_mass_velocity().numpy(), g['center_of_mass_velocity'], **tol)
    np.testing.assert_allclose(a0.get_angular_momentum().numpy(), g['angular_momentum'], **tol)
    Ip, basis = a0.get_moments_of_inertia(vectors=True)
    np.testing.assert_allclose(Ip.numpy(), g['moments_of_inertia'], rtol=1e-10)
    np.testing.assert_allclose(a0.get_moments_of_inertia().numpy(), g['moments_of_inertia'], rtol=1e-10)
    # principal axes: rows, defined up to a sign each
    np.testing.assert_allclose(np.abs((basis.numpy() * g['principal_axes']).sum(1)), 1.0, rtol=1e-9)
    cases = {'zero_translation': lambda a: md.zero_translation(a),
             'zero_translation_keep_energy': lambda a: md.zero_translation(a, preserve_temperature=False),
             'zero_rotation': lambda a: md.zero_rotation(a),
             'zero_rotation_keep_energy': lambda a: md.zero_rotation(a, preserve_temperature=False),
             'both': lambda a: md.zero_rotation(md.zero_translation(a)),
             'scale_momenta': lambda a: md.scale_momenta(a, 300.0 * md.KB_EV)}
    for tag, fn in cases.items():
        a = fn(_utils_atoms(g))
        np.testing.assert_allclose(a.get_momenta().numpy(), g[f'{tag}/p'], rtol=1e-10, atol=1e-13, err_msg=tag)
        np.testing.assert_allclose(float(a.get_temperature()), g[f'{tag}/temperature'], rtol=1e-11, err_msg=tag)
    np.testing.assert_allclose(a0.get_positions().numpy(), g['R'])          # the inputs are untouched


def test_zero_rotation_and_zero_translation():
    """The properties the reference's tests assert (tests/test_mdx_utils.py:6-88, there against ASE's Stationary /
    ZeroRotation): no centre-of-mass velocity, no angular momentum, temperature preserved; and test_momenta_scaling."""
    g = golden('reference_mdx_utils.npz')
    a = _utils_atoms(g)
    T0 = float(a.get_temperature())
    assert np.abs(a.get_center_of_mass_velocity().numpy()).max() > 1e-4 and np.abs(a.get_angular_momentum().numpy()).max() > 1e-2
    a = md.zero_translation(a)
    assert np.abs(a.get_center_of_mass_velocity().numpy()).max() < 1e-12
    assert abs(float(a.get_temperature()) - T0) < 1e-12
    a = md.zero_rotation(a)
    assert np.abs(a.get_angular_momentum().numpy()).max() < 1e-12
    assert np.abs(a.get_center_of_mass_velocity().numpy()).max() < 1e-12     # a rotation about the centre of mass keeps it
    assert abs(float(a.get_temperature()) - T0) < 1e-12
    a = md.scale_momenta(a, 500.0 * md.KB_EV)
    assert abs(float(a.get_temperature()) / md.KB_EV - 500.0) < 1e-9
    # a linear molecule (one vanishing moment of inertia) is handled: that axis is skipped
    lin = md.AtomsX(positions=torch.tensor([[0.0, 0, 0], [0, 0, 1.1], [0, 0, 2.3]], dtype=torch.float64),
                    momenta=torch.tensor([[0.1, 0, 0], [0, 0.05, 0], [-0.1, 0.02, 0.3]], dtype=torch.float64),
                    masses=torch.tensor([15.999, 12.011, 15.999], dtype=torch.float64), numbers=torch.tensor([8, 6, 8]))
    out = md.zero_rotation(lin, preserve_temperature=False)
    assert torch.isfinite(out.get_momenta()).all()
    L = out.get_angular_momentum().numpy()
    assert np.abs(L[:2]).max() < 1e-12


@pytest.mark.parametrize('tag', ['baoab', 'baoab_free'])
def test_baoab_langevin_matches_reference_source(monkeypatch, tag):
    """BAOABLangevinX (mlff/mdx/integrator.py:162-233) executed from the reference's source on the reference's AtomsX with
    its jax.random draws recorded (tests/golden/make_reference_mdx_utils_golden.py); mlff_b200.md.BAOABLangevinX fed the same
    numbers reproduces the 30-step trajectory -- with the centre-of-mass and rotation fixes of every step, and without."""
    g = golden('reference_mdx_utils.npz')
    D0, K = float(g['D0']), float(g['K'])
    stream = iter(g[f'{tag}/noise'])
    monkeypatch.setattr(md.torch, 'randn', lambda *shape, **kw: torch.as_tensor(next(stream)))

    class Springs:
        def __call__(self, atoms):
            if atoms.cache is None:
                R = atoms.get_positions()
                d = R[None, :, :] - R[:, None, :]
                eye = torch.eye(len(R), dtype=R.dtype)
                r = d.norm(dim=-1) + eye
                f = K * (r - D0) / r * (1 - eye)
                atoms.cache = {'energy': 0.25 * K * ((r - D0) ** 2 * (1 - eye)).sum(), 'forces': (f[:, :, None] * d).sum(1)}
            return atoms.cache

    class NoList:
        def update(self, positions):
            return self

    atoms = md.AtomsX(positions=torch.as_tensor(g['baoab_R0']), momenta=torch.as_tensor(g['baoab_p0']),
                      masses=torch.as_tensor(g['baoab_masses']), numbers=torch.ones(5, dtype=torch.int32), neighbors=NoList())
    fix = tag == 'baoab'
    integrator = md.BAOABLangevinX.create(float(g['baoab_dt']), float(g['baoab_temperature']), Springs(),
                                          gamma=float(g['baoab_gamma']), fixcm=fix, fixrot=fix)
    for k in range(len(g[f'{tag}/E'])):
        integrator, atoms, prop = integrator.step(atoms)
        np.testing.assert_allclose(atoms.get_positions().numpy(), g[f'{tag}/R'][k], rtol=1e-10, atol=1e-12, err_msg=str(k))
        np.testing.assert_allclose(atoms.get_momenta().numpy(), g[f'{tag}/p'][k], rtol=1e-10, atol=1e-12, err_msg=str(k))
        assert abs(float(prop['energy']) - g[f'{tag}/E'][k]) < 1e-10
    if fix:
        assert np.abs(atoms.get_angular_momentum().numpy()).max() < 1e-6      # pair forces exert no net torque


def test_maxwell_boltzmann_distribution():
    """mlff/mdx/distributions.py:5-20: p = xi sqrt(m T0); the sample temperature (3n degrees of freedom) approaches T0."""
    masses = torch.full((20000,), 12.011, dtype=torch.float64)
    T0 = 300.0 * md.KB_EV
    p = md.maxwell_boltzmann(masses, T0, seed=3)
    a = md.AtomsX(positions=torch.zeros(20000, 3, dtype=torch.float64), momenta=p, masses=masses,
                  numbers=torch.full((20000,), 6, dtype=torch.int32))
    assert abs(float(a.get_temperature(eckart_frame=False)) / T0 - 1.0) < 0.02
    assert torch.equal(p, md.maxwell_boltzmann(masses, T0, seed=3)) and not torch.equal(p, md.maxwell_boltzmann(masses, T0, seed=4))


def test_calculator_with_stress_property():
    """CalculatorX.create(engine, implemented_properties=('energy', 'forces', 'stress')) -- mlff/mdx/calculator.py:16-52:
    the extra entry is dE/d(strain) of the periodic cell, here checked against central differences of the calculator's own
    energy under r -> (1 + eps) r, cell -> cell (1 + eps)^T."""
    sol = golden('solid_0.npz')
    cfg = o.OracleConfig(F=16, n_layers=1, degrees=(1, 2), n_heads=2)
    p = o.init_params(cfg, 2, embed_scale=4.0)
    eng = OracleEngine(cfg, p)
    R, z, cell = sol['R'].astype(np.float64), sol['z'].astype(np.int32), sol['unit_cell'].astype(np.float64)

    def evaluate(Rx, cx, props):
        n = len(z)
        atoms = md.AtomsX(positions=torch.as_tensor(Rx), momenta=torch.zeros(n, 3, dtype=torch.float64),
                          masses=torch.ones(n, dtype=torch.float64), numbers=torch.as_tensor(z), cell=cx,
                          pbc=(True, True, True)).init_spatial_partitioning(eng, cfg.r_cut, 0.0)
        return md.CalculatorX.create(eng, implemented_properties=props)(atoms)

    out = evaluate(R, cell, ('energy', 'forces', 'stress'))
    assert set(out) == {'energy', 'forces', 'stress'} and out['stress'].shape == (3, 3)
    assert set(evaluate(R, cell, ('energy', 'forces'))) == {'energy', 'forces'}
    # the calculator hands float32 positions / cell to the engine: finite differences see that rounding, so the step is large
    h = 2e-3
    for a, b in ((0, 0), (1, 2), (2, 2)):
        vals = []
        for sgn in (1.0, -1.0):
            eps = np.zeros((3, 3))
            eps[a, b] += 0.5 * sgn * h
            eps[b, a] += 0.5 * sgn * h
            vals.append(float(evaluate(R + R @ eps.T, cell + cell @ eps.T, ('energy', 'forces'))['energy']))
        fd = (vals[0] - vals[1]) / (2 * h)
        sym = 0.5 * float(out['stress'][a, b] + out['stress'][b, a])
        assert abs(fd - sym) <= 2e-3 * max(1.0, abs(fd)), (a, b, fd, sym)
    ii, jj, S = o.primitive_neighbor_list(R, cell, (True, True, True), cfg.r_cut)
    _, _, st = o.energy_forces_stress(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    np.testing.assert_allclose(out['stress'].numpy(), st.numpy(), rtol=1e-4, atol=1e-4 * float(st.abs().max()))
    with pytest.raises(NotImplementedError):
        md.CalculatorX.create(eng, implemented_properties=('energy', 'forces', 'heat_flux'))


def test_ase_style_calculator_glue_with_a_stand_in_engine(monkeypatch):
    """mlffCalculator.calculate (mlff/md/calculator.py:121-147) host logic on the CPU: the single read-back is sliced into
    energy, forces and -- with calculate_stress -- the strain derivative; overflow re-allocates; no cell, no stress entry.
    The engine is the float64 oracle stand-in (the CUDA engine is exercised by tests/test_gpu_parity.py)."""
    from mlff_b200 import calculator as C
    from mlff_b200.config import So3kratesConfig
    sol = golden('solid_0.npz')
    cfg = o.OracleConfig(F=16, n_layers=1, degrees=(1, 2), n_heads=2)
    p = o.init_params(cfg, 2, embed_scale=4.0)

    class StandIn(OracleEngine):
        def __init__(self, pc, device, flags):
            super().__init__(cfg, p)
            self.device = torch.device('cpu')
            self._status = torch.zeros(1, dtype=torch.int32)

        def load_params(self, flat):
            pass

        def raise_for_status(self, s):
            assert s == 0
            return 0

    monkeypatch.setattr(C, 'So3kEngine', StandIn)
    monkeypatch.setattr(C.mlffCalculator, '_pinned', lambda self, name, arr: torch.from_numpy(arr))
    pc = So3kratesConfig(F=cfg.F, n_layer=cfg.n_layers, degrees=tuple(cfg.degrees), n_heads=cfg.n_heads)
    R, z, cell = sol['R'].astype(np.float64), sol['z'].astype(np.int32), sol['unit_cell'].astype(np.float64)
    ii, jj, S = o.primitive_neighbor_list(R, cell, (True, True, True), cfg.r_cut)
    Eo, Fo, So = o.energy_forces_stress(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    calc = C.mlffCalculator(pc, {'embed/embedding': None}, calculate_stress=True, energy_scale=2.0)
    calc.capacity = 500                                              # overflow branch first
    res = calc.calculate(R, z, cell=cell, pbc=(True, True, True))
    assert calc.n_edges == len(ii) and calc.capacity >= len(ii)
    assert abs(res['energy'] - 2.0 * float(Eo)) <= 1e-6 * abs(float(Eo))
    np.testing.assert_allclose(res['forces'], 2.0 * Fo.numpy(), atol=2e-5)
    np.testing.assert_allclose(calc.get_stress(), 2.0 * So.numpy(), rtol=1e-4, atol=1e-4 * float(So.abs().max()))
    assert res['forces'].shape == (len(z), 3) and res['stress'].shape == (3, 3)
    plain = C.mlffCalculator(pc, {'embed/embedding': None}).calculate(R, z, cell=cell, pbc=(True, True, True))
    assert set(plain) == {'energy', 'forces'} and abs(plain['energy'] - float(Eo)) <= 1e-6 * abs(float(Eo))
    assert 'stress' not in calc.calculate(R, z)                     # non-periodic call
