"""Golden values from the REFERENCE'S OWN mdx utilities: mlff/mdx/utils.py (zero_rotation :6-36, zero_translation :39-67,
scale_momenta :70-83) and the AtomsX accessors they use (mlff/mdx/atoms.py: get_moments_of_inertia :248-288,
get_center_of_mass :385-400, get_center_of_mass_velocity :420-428, get_angular_momentum :430-441, get_temperature :211-226),
executed unmodified from /root/reference on the numpy stand-ins of ref_shim.py.  jax arrays are immutable -- the
reference writes `pos -= com` on what `get_positions()` returns -- so the arrays handed to AtomsX are an ndarray
subclass whose in-place operators return new arrays.  The system is the ethanol frame the reference's own
tests/test_mdx_utils.py uses (frame 100) with seeded Maxwell-Boltzmann momenta at 500 K.
tests/test_md.py::test_momentum_utilities_match_reference_source replays the same state through mlff_b200.md.
Also a 30-step BAOABLangevinX trajectory (mlff/mdx/integrator.py:162-233; it calls the utilities on every step) with its
noise recorded, for tests/test_md.py::test_baoab_langevin_matches_reference_source.

    python tests/golden/make_reference_mdx_utils_golden.py          (build container only; needs /root/reference)
"""
import dataclasses
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as rs  # noqa: E402

KB = 8.617333262e-5            # eV / K (ase.units.kB)


class Immutable(np.ndarray):
    """numpy array with jax's semantics for augmented assignment: `a -= b` rebinds, it never writes through."""

    def __isub__(self, o): return np.subtract(self, o).view(Immutable)
    def __iadd__(self, o): return np.add(self, o).view(Immutable)
    def __imul__(self, o): return np.multiply(self, o).view(Immutable)
    def __itruediv__(self, o): return np.true_divide(self, o).view(Immutable)


def imm(a):
    return np.array(a, dtype=np.float64).view(Immutable)


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
    ase, ase_atoms = types.ModuleType('ase'), types.ModuleType('ase.atoms')
    ase_atoms.Atoms = type('Atoms', (), {})
    ase.atoms = ase_atoms
    sys.modules.update({'ase': ase, 'ase.atoms': ase_atoms})
    mdx = types.ModuleType('mlff.mdx')
    mdx.__path__ = []
    sys.modules['mlff.mdx'] = mdx
    utils_pkg = types.ModuleType('mlff.utils')
    utils_pkg.__path__ = []
    sys.modules['mlff.utils'] = utils_pkg
    st = rs.load('mlff.utils.structures')
    rs.export('mlff.utils', st, 'Graph', 'Neighbors', 'System')
    atoms_mod = rs.load('mlff.mdx.atoms')
    utils = rs.load('mlff.mdx.utils')

    d = np.load('/root/reference/tests/test_data/ethanol.npz')
    R, z = d['R'][100].astype(np.float64), d['z'].astype(np.int32)
    masses = np.array([{1: 1.008, 6: 12.011, 8: 15.999}[int(a)] for a in z])
    rng = np.random.default_rng(1)
    p0 = rng.normal(size=R.shape) * np.sqrt(masses * KB * 500.0)[:, None]        # Maxwell-Boltzmann at 500 K

    def make(p):
        return atoms_mod.AtomsX(numbers=z, masses=imm(masses), pbc=np.zeros(3, bool), positions=imm(R), momenta=imm(p))

    a0 = make(p0)
    Ip, basis = a0.get_moments_of_inertia(vectors=True)
    out = {'R': R, 'z': z, 'masses': masses, 'p0': p0,
           'temperature': np.array(a0.get_temperature()), 'kinetic_energy': np.array(a0.get_kinetic_energy()),
           'center_of_mass': np.array(a0.get_center_of_mass()),
           'center_of_mass_velocity': np.array(a0.get_center_of_mass_velocity()),
           'angular_momentum': np.array(a0.get_angular_momentum()),
           'moments_of_inertia': np.array(Ip), 'principal_axes': np.array(basis)}
    cases = {'zero_translation': lambda a: utils.zero_translation(a),
             'zero_translation_keep_energy': lambda a: utils.zero_translation(a, preserve_temperature=False),
             'zero_rotation': lambda a: utils.zero_rotation(a),
             'zero_rotation_keep_energy': lambda a: utils.zero_rotation(a, preserve_temperature=False),
             'both': lambda a: utils.zero_rotation(utils.zero_translation(a)),
             'scale_momenta': lambda a: utils.scale_momenta(a, T0=300.0 * KB)}
    for tag, fn in cases.items():
        a = fn(make(p0))
        out[f'{tag}/p'] = np.array(a.get_momenta())
        out[f'{tag}/temperature'] = np.array(a.get_temperature())
        out[f'{tag}/angular_momentum'] = np.array(a.get_angular_momentum())
        out[f'{tag}/center_of_mass_velocity'] = np.array(a.get_center_of_mass_velocity())
    assert np.allclose(np.asarray(make(p0).get_positions()), R)                   # nothing wrote through

    # ---- BAOABLangevinX (mlff/mdx/integrator.py:162-233), which calls the two utilities on every step: executed from the
    # reference's source on the reference's AtomsX, its jax.random draws recorded (note Normal.sample: the variance has
    # shape (n, 1), so ONE normal number per atom is broadcast over x, y, z -- mirrored as is)
    calc_mod = types.ModuleType('mlff.mdx.calculator')
    calc_mod.CalculatorX = None
    sys.modules['mlff.mdx.calculator'] = calc_mod
    noise_rng, noise = np.random.default_rng(17), []

    def normal(key=None, shape=None, dtype=None):
        noise.append(noise_rng.normal(size=shape))
        return noise[-1].copy()

    sys.modules['jax'].random = types.SimpleNamespace(PRNGKey=lambda s: s, normal=normal,
                                                      split=lambda key, num=2: tuple(range(num)))
    integ = rs.load('mlff.mdx.integrator')
    D0, K = 1.2, 5.0

    def calculator(atoms):                      # the analytic spring network of make_reference_md_golden.py
        Rc = np.asarray(atoms.get_positions())
        dd = Rc[None, :, :] - Rc[:, None, :]
        r = np.linalg.norm(dd, axis=-1) + np.eye(len(Rc))
        f = K * (r - D0) / r
        np.fill_diagonal(f, 0.0)
        return {'energy': 0.25 * K * ((r - D0) ** 2 * (1 - np.eye(len(Rc)))).sum(), 'forces': imm((f[:, :, None] * dd).sum(1))}

    Rb = np.array([[0.0, 0.0, 0.0], [1.5, 0.1, 0.0], [0.6, 1.3, 0.2], [0.5, 0.4, 1.6], [1.4, 1.2, 1.1]])
    mb = np.array([1.008, 12.011, 15.999, 1.008, 14.007])
    pb = np.random.default_rng(0).normal(0.0, 0.3, Rb.shape) * np.sqrt(mb)[:, None]
    part = atoms_mod.SpatialPartitioning(allocate_fn=None, update_fn=lambda a, nbrs: nbrs, cutoff=0.0, skin=0.0,
                                         capacity_multiplier=1.0)
    for tag, kw in (('baoab', dict(fixcm=True, fixrot=True)), ('baoab_free', dict(fixcm=False, fixrot=False))):
        noise.clear()
        atoms = atoms_mod.AtomsX(numbers=np.ones(5, np.int32), masses=imm(mb), pbc=np.zeros(3, bool), positions=imm(Rb),
                                 momenta=imm(pb), spatial_partitioning=part)
        integrator = integ.BAOABLangevinX.create(timestep=0.02, temperature=0.05, calculator=calculator, gamma=0.7, **kw)
        Rs, ps, Es = [], [], []
        for _ in range(30):
            integrator, atoms, prop = integrator.step(atoms)
            Rs.append(np.array(atoms.get_positions()))
            ps.append(np.array(atoms.get_momenta()))
            Es.append(float(prop['energy']))
        out[f'{tag}/R'], out[f'{tag}/p'], out[f'{tag}/E'], out[f'{tag}/noise'] = map(np.array, (Rs, ps, Es, noise))
    out.update({'baoab_R0': Rb, 'baoab_p0': pb, 'baoab_masses': mb, 'baoab_dt': np.array(0.02), 'baoab_temperature': np.array(0.05),
                'baoab_gamma': np.array(0.7), 'D0': np.array(D0), 'K': np.array(K)})
    np.savez_compressed(os.path.join(HERE, 'reference_mdx_utils.npz'), **out)
    print('wrote reference_mdx_utils.npz; L0 =', out['angular_momentum'], '-> after zero_rotation',
          out['zero_rotation/angular_momentum'], ' v_com after zero_translation', out['zero_translation/center_of_mass_velocity'])


if __name__ == '__main__':
    main()
