"""Golden MD trajectories from the REFERENCE'S OWN INTEGRATOR CODE (mlff/mdx/integrator.py: VelocityVerletX :312-342,
NoseHooverX :16-83, BerendsenX :237-310, LangevinX :86-159 with its noise recorded), executed unmodified from /root/reference on the numpy stand-ins of ref_shim.py (`flax.struct`
becomes a plain dataclass with `.replace`).  The integrators are duck-typed on the atoms object and the calculator, so
both are small numpy objects here: the atoms object implements exactly the accessors the reference's AtomsX offers
(mlff/mdx/atoms.py:174-246, 365-383) and the calculator is an analytic spring network whose forces the test
re-implements in torch.  tests/test_md.py::test_integrators_match_reference_source replays the same initial state
through mlff_b200.md and compares positions, momenta and the thermostat variable step by step.

    python tests/golden/make_reference_md_golden.py              (build container only; needs /root/reference)
"""
import dataclasses
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as rs  # noqa: E402

D0, K = 1.2, 5.0


def spring_forces(R):
    """E = sum_{i<j} K/2 (|R_j - R_i| - D0)^2"""
    d = R[None, :, :] - R[:, None, :]
    r = np.linalg.norm(d, axis=-1) + np.eye(len(R))
    f = K * (r - D0) / r
    np.fill_diagonal(f, 0.0)
    E = 0.25 * K * ((r - D0) ** 2 * (1 - np.eye(len(R)))).sum()
    return E, (f[:, :, None] * d).sum(1)


@dataclasses.dataclass
class NumpyAtoms:
    positions: np.ndarray
    momenta: np.ndarray
    masses: np.ndarray

    def get_positions(self): return self.positions
    def get_momenta(self): return self.momenta.copy()          # jax arrays are immutable: `p += ...` must not alias
    def get_masses(self): return self.masses
    def get_number_of_atoms(self): return len(self.masses)
    def get_atomic_numbers(self): return np.ones(len(self.masses), np.int32)
    def get_velocities(self): return self.momenta / self.masses[:, None]
    def get_kinetic_energy(self): return 0.5 * (self.momenta * self.get_velocities()).sum()
    def get_temperature(self): return 2 * self.get_kinetic_energy() / (3 * len(self.masses) - 6)   # atoms.py:210-225
    def update_positions(self, x, update_neighbors=True): return dataclasses.replace(self, positions=x)
    def update_momenta(self, p): return dataclasses.replace(self, momenta=p)
    def update_velocities(self, v): return dataclasses.replace(self, momenta=v * self.masses[:, None])


def calculator(atoms):
    E, F = spring_forces(atoms.get_positions())
    return {'energy': E, 'forces': F}


def main():
    rs.install()
    struct = types.ModuleType('flax.struct')

    def struct_dataclass(cls):
        cls = dataclasses.dataclass(cls)
        cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
        return cls

    struct.dataclass = struct_dataclass
    struct.field = lambda pytree_node=True, **kw: dataclasses.field(**kw)
    sys.modules['flax.struct'] = struct
    sys.modules['flax'].struct = struct
    for name, attrs in (('mlff.mdx', ()), ('mlff.mdx.atoms', ('AtomsX',)), ('mlff.mdx.calculator', ('CalculatorX',)),
                        ('mlff.mdx.utils', ('zero_rotation', 'zero_translation'))):
        m = types.ModuleType(name)
        m.__path__ = []
        for a in attrs:
            setattr(m, a, None)
        sys.modules[name] = m
    # LangevinX draws xi, eta from jax.random: here from a numpy stream that is RECORDED, so the test can feed the same
    # numbers to mlff_b200.md.LangevinX and compare the deterministic part of the update
    noise_rng, noise = np.random.default_rng(42), []

    def normal(key=None, shape=None, dtype=None):
        noise.append(noise_rng.normal(size=shape))
        return noise[-1].copy()

    sys.modules['jax'].random = types.SimpleNamespace(PRNGKey=lambda s: s, normal=normal,
                                                      split=lambda key, num=2: tuple(range(num)))
    integ = rs.load('mlff.mdx.integrator')

    rng = np.random.default_rng(0)
    R0 = np.array([[0.0, 0.0, 0.0], [1.5, 0.1, 0.0], [0.6, 1.3, 0.2], [0.5, 0.4, 1.6], [1.4, 1.2, 1.1]])
    masses = np.array([1.008, 12.011, 15.999, 1.008, 14.007])
    p0 = rng.normal(0.0, 0.3, R0.shape) * np.sqrt(masses)[:, None]
    out = {'R0': R0, 'p0': p0, 'masses': masses, 'D0': np.array(D0), 'K': np.array(K)}

    def record(tag, integrator, steps):
        atoms = NumpyAtoms(R0.copy(), p0.copy(), masses)
        Rs, ps, zs, Es = [], [], [], []
        for _ in range(steps):
            integrator, atoms, prop = integrator.step(atoms)
            Rs.append(atoms.get_positions().copy())
            ps.append(atoms.get_momenta().copy())
            zs.append(float(getattr(integrator, 'zeta', 0.0)))
            Es.append(float(prop['energy']))
        out[f'{tag}/R'], out[f'{tag}/p'], out[f'{tag}/zeta'], out[f'{tag}/E'] = map(np.array, (Rs, ps, zs, Es))

    dt = 0.02
    out['dt'] = np.array(dt)
    record('velocity_verlet', integ.VelocityVerletX.create(timestep=dt, calculator=calculator), 50)
    out['nh_temperature'], out['nh_ttime'] = np.array(0.05), np.array(10.0)
    record('nose_hoover', integ.NoseHooverX.create(timestep=dt, temperature=0.05, ttime=10.0, calculator=calculator), 50)
    out['be_temperature'], out['be_time_constant'] = np.array(0.05), np.array(0.5)
    record('berendsen', integ.BerendsenX.create(timestep=dt, temperature=0.05, calculator=calculator, time_constant=0.5), 50)
    out['lv_temperature'], out['lv_friction'] = np.array(0.05), np.array(0.8)
    record('langevin', integ.LangevinX.create(timestep=dt, temperature=0.05, friction=0.8, calculator=calculator), 30)
    out['langevin/noise'] = np.array(noise)                       # (2 * steps, n, 3): xi, eta, xi, eta, ...
    np.savez_compressed(os.path.join(HERE, 'reference_md.npz'), **out)
    print('wrote reference_md.npz; final VV E =', out['velocity_verlet/E'][-1], ' NH zeta =', out['nose_hoover/zeta'][-1])


if __name__ == '__main__':
    main()
