"""MD inner loop around the force call, state resident on the device.

Mirrors the pieces of the reference's `mdx` package that sit immediately around the hot path, with the same names and
update rules, on torch tensors instead of jax arrays:

* `AtomsX`            mlff/mdx/atoms.py (positions, momenta, masses, numbers, cell, neighbours; get_* accessors :202-246,
                      `update_positions(..., update_neighbors=True)` :365-383)
* `SkinNeighborList`  mlff/mdx/atoms.py:488-510 -> glp `quadratic_neighbor_list(cell, cutoff, skin, capacity_multiplier)`:
                      the list is built with cutoff + skin at reference positions and re-used until an atom has moved
                      more than skin / 2; edges between r_cut and r_cut + skin contribute exactly zero (cosine cutoff)
* `CalculatorX`       mlff/mdx/calculator.py:46-49 (energy + forces of an AtomsX)
* `VelocityVerletX`   mlff/mdx/integrator.py:313-342 (the reference evaluates the forces twice per step; the second
                      evaluation of a step is the first of the next one, so it is cached here -- same trajectory)
* `NoseHooverX`, `LangevinX`, `BerendsenX`   mlff/mdx/integrator.py:17-83, 86-159, 237-310
* `simulate`          mlff/mdx/simulate.py:60-82 (kinetic / potential energy per step, one host read-back at the end)
* `GradientDescent`, `LBFGS`   mlff/mdx/optimizer.py:16-52, 55-149 (structure relaxation; same `create` / `step` /
                      `minimize(atoms, max_steps, tol)` contract and error behaviour)

Everything per step is O(N) tensor arithmetic on the device plus the engine calls; the only host synchronisation is the
scalar "has any atom moved more than skin / 2" decision of the neighbour list (and the list size after a rebuild).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, Optional, Sequence

import numpy as np
import torch

KB_EV = 8.617333262e-5            # Boltzmann constant in eV / K (ase.units.kB)
FS = 0.09822694788464063          # one femtosecond in ASE time units (ase.units.fs)


class SkinNeighborList:
    """Neighbour list with a skin, re-used while no atom has moved more than skin / 2 since it was built."""

    def __init__(self, engine, cutoff: float, skin: float, cell: Optional[np.ndarray] = None,
                 pbc: Sequence[bool] = (False, False, False), capacity_multiplier: float = 1.25, mode: Optional[int] = None):
        self.engine, self.cutoff, self.skin = engine, float(cutoff), float(skin)
        self.cell = None if cell is None else np.asarray(cell, dtype=np.float64).reshape(3, 3)
        self.pbc = tuple(bool(b) for b in pbc)
        self.capacity_multiplier = capacity_multiplier
        self.mode = mode
        self.capacity = 0
        self.reference_positions: Optional[torch.Tensor] = None
        self.idx_i = self.idx_j = self.shifts = None
        self.count = 0
        self.n_builds = 0

    def allocate(self, positions: torch.Tensor) -> 'SkinNeighborList':
        """Build the list at `positions` (glp `allocate`): sizes the capacity from the actual count."""
        kw = {} if self.mode is None else {'mode': self.mode}
        cap = self.capacity if self.capacity > 0 else 64 * int(positions.shape[0]) + 16
        while True:
            ii, jj, S, count = self.engine.neighbor_list(positions, self.cutoff + self.skin, cap, cell=self.cell,
                                                         pbc=self.pbc, **kw)
            # edge count + device status word in one read: errors of the evaluations since the last build (asymmetric
            # list, bad Z, tensor-kernel time-out) surface here, once per rebuild
            n = self.engine.count_and_status(count) if hasattr(self.engine, 'count_and_status') else int(count.item())
            if n <= cap:
                break
            cap = int(self.capacity_multiplier * n) + 16          # overflow: re-allocate (indices.py:264-272 semantics)
        if self.capacity == 0 or cap > self.capacity:
            self.capacity = max(int(self.capacity_multiplier * n) + 16, n)
        # full-capacity arrays (padded with -1 / 0 behind `count`): their shape only changes when the capacity grows, which
        # lets the calculator replay one CUDA graph between re-allocations
        self.idx_i, self.idx_j, self.shifts, self.count = ii, jj, S, n
        self.reference_positions = positions.clone()
        self.n_builds += 1
        return self

    def update(self, positions: torch.Tensor) -> 'SkinNeighborList':
        """glp `update`: rebuild only if some atom moved more than skin / 2 since the last build."""
        if self.reference_positions is None:
            return self.allocate(positions)
        d2 = ((positions - self.reference_positions) ** 2).sum(-1).max()
        if float(d2.item()) > (0.5 * self.skin) ** 2:
            self.allocate(positions)
        return self


@dataclass
class AtomsX:
    positions: torch.Tensor                     # (N,3) float64
    momenta: torch.Tensor                       # (N,3) float64
    masses: torch.Tensor                        # (N)   float64
    numbers: torch.Tensor                       # (N)   int32
    cell: Optional[np.ndarray] = None           # (3,3) row-wise (host)
    pbc: Sequence[bool] = (False, False, False)
    neighbors: Optional[object] = None          # SkinNeighborList, or DDNeighbors (domain decomposition)
    cache: Optional[Dict[str, torch.Tensor]] = None     # properties evaluated at `positions` (force re-use)

    def get_positions(self): return self.positions
    def get_momenta(self): return self.momenta
    def get_masses(self): return self.masses
    def get_number_of_atoms(self): return int(self.numbers.shape[0])
    def get_velocities(self): return self.momenta / self.masses[:, None]

    def get_kinetic_energy(self):
        return 0.5 * (self.momenta * self.get_velocities()).sum()

    def get_temperature(self, eckart_frame: bool = True):
        """Temperature in eV (mlff/mdx/atoms.py:211-226): 2 E_kin / DOF, DOF = 3n - 6 in the Eckart frame."""
        n = self.get_number_of_atoms()
        dof = 3 * n - 6 if eckart_frame else 3 * n
        return 2.0 * self.get_kinetic_energy() / dof

    def get_center_of_mass(self):
        """mlff/mdx/atoms.py:385-400 (Cartesian; periodic boundary conditions ignored)."""
        return (self.masses[:, None] * self.positions).sum(0) / self.masses.sum()

    def get_center_of_mass_velocity(self):
        """mlff/mdx/atoms.py:420-428."""
        return self.momenta.sum(0) / self.masses.sum()

    def get_angular_momentum(self):
        """Total angular momentum about the centre of mass, mlff/mdx/atoms.py:430-441."""
        return torch.linalg.cross(self.positions - self.get_center_of_mass(), self.momenta).sum(0)

    def get_moments_of_inertia(self, vectors: bool = False):
        """Principal moments of the inertia tensor about the centre of mass (ascending) and, with vectors=True, the
        principal axes as ROWS -- mlff/mdx/atoms.py:248-288."""
        r = self.positions - self.get_center_of_mass()
        m = self.masses
        eye = torch.eye(3, dtype=r.dtype, device=r.device)
        inertia = (m * (r * r).sum(-1)).sum() * eye - torch.einsum('a,ai,aj->ij', m, r, r)
        evals, evecs = torch.linalg.eigh(inertia)
        return (evals, evecs.T) if vectors else evals

    def update_positions(self, positions: torch.Tensor, update_neighbors: bool = True) -> 'AtomsX':
        if update_neighbors:
            if self.neighbors is None:
                raise RuntimeError('AtomsX object has no spatial partitioning.')
            self.neighbors.update(positions)
        return replace(self, positions=positions, cache=None)

    def update_momenta(self, momenta: torch.Tensor) -> 'AtomsX':
        return replace(self, momenta=momenta)

    def update_velocities(self, velocities: torch.Tensor) -> 'AtomsX':
        return replace(self, momenta=velocities * self.masses[:, None])      # mlff/mdx/atoms.py:174-186

    def init_domain_decomposition(self, session) -> 'AtomsX':
        """The lists of a domain-decomposed run live in `session` (mlff_b200.dist.DDSession); use with CalculatorDD."""
        return replace(self, neighbors=DDNeighbors(session), cache=None)

    def init_spatial_partitioning(self, engine, cutoff: float, skin: float, capacity_multiplier: float = 1.25,
                                  mode: Optional[int] = None) -> 'AtomsX':
        nl = SkinNeighborList(engine, cutoff, skin, self.cell, self.pbc, capacity_multiplier, mode).allocate(self.positions)
        return replace(self, neighbors=nl, cache=None)


class CalculatorX:
    """energy + forces (+ stress) of an AtomsX through the engine (one structure; the skin list of the atoms object).
    implemented_properties as in mlff/mdx/calculator.py:16-52: with 'stress' the result also holds dE/d(strain) (3,3) of the
    periodic cell -- what the reference differentiates through glp's strain_graph (edge vectors move as r -> (1 + eps) r;
    not divided by the volume)."""

    def __init__(self, engine, energy_scale: float = 1.0, cuda_graph: bool = False,
                 implemented_properties: Sequence[str] = ('energy', 'forces')):
        """cuda_graph=True replays one captured graph per evaluation while the list capacity is unchanged (small,
        launch-bound systems); the engine must provide `graphed_energy_forces`."""
        self.engine = engine
        self.energy_scale = energy_scale
        self.n_calls = 0
        self.cuda_graph = cuda_graph
        self._graphed = None
        self.implemented_properties = tuple(implemented_properties)
        unknown = set(self.implemented_properties) - {'energy', 'forces', 'stress'}
        if unknown:
            raise NotImplementedError(f'properties {sorted(unknown)} are not implemented')
        if 'stress' in self.implemented_properties and cuda_graph:
            raise NotImplementedError('the stress is evaluated by a second call on the workspace: use cuda_graph=False')

    @classmethod
    def create(cls, engine, implemented_properties: Sequence[str] = ('energy', 'forces'), energy_scale: float = 1.0):
        return cls(engine, energy_scale, False, implemented_properties)

    def __call__(self, atoms: AtomsX) -> Dict[str, torch.Tensor]:
        if atoms.cache is not None:
            return atoms.cache
        nl = atoms.neighbors
        periodic = atoms.cell is not None
        R32 = atoms.positions.to(torch.float32)
        cell32 = None
        if periodic:
            cell32 = torch.as_tensor(np.asarray(atoms.cell, dtype=np.float32), device=R32.device)[None]
        args = (R32[None], atoms.numbers[None], nl.idx_i[None], nl.idx_j[None], cell32, nl.shifts[None] if periodic else None)
        if self.cuda_graph:
            shape = (1, int(R32.shape[0]), int(nl.idx_i.shape[0]))
            if self._graphed is None or self._graphed.shape != shape:
                self._graphed = self.engine.graphed_energy_forces(*shape, periodic)
            E, F = self._graphed(*args)
        elif 'stress' in self.implemented_properties:
            if not periodic:
                raise ValueError('stress needs a periodic cell')
            E, F, _, stress = self.engine.energy_forces(*args, want_stress=True)
        else:
            E, F, _ = self.engine.energy_forces(*args)
        self.n_calls += 1
        out = {'energy': self.energy_scale * E.reshape(()).to(torch.float64),
               'forces': self.energy_scale * F[0].to(torch.float64)}
        if 'stress' in self.implemented_properties:
            out['stress'] = self.energy_scale * stress[0].to(torch.float64)
        atoms.cache = out
        return out


class DDNeighbors:
    """What the integrators ask of `atoms.neighbors` when the lists live in a domain-decomposition session
    (mlff_b200.dist.DDSession keeps the per-rank plan within its own skin): `update` is a no-op."""

    def __init__(self, session):
        self.session = session

    def update(self, positions: torch.Tensor) -> 'DDNeighbors':
        return self

    @property
    def n_builds(self) -> int:
        return self.session.n_builds

    @property
    def count(self) -> int:
        return int(self.session.dd.P) if getattr(self.session.dd, 'P', None) is not None else 0


class CalculatorDD:
    """energy + forces of ONE large periodic system decomposed over the ranks of a process group (spatial bricks, halo
    exchange per layer: mlff_b200.dist.DDSession).  Every rank holds the whole AtomsX -- positions replicated, the O(N)
    integrator update replicated and deterministic -- evaluates its brick, and the owned forces and energies are summed
    into the full arrays with ONE all-reduce of 3 N + 1 float64 numbers (NCCL; skipped for a single rank).  The reference's
    mdx loop is single-device; this is how its inner loop scales past one GPU."""

    def __init__(self, session, energy_scale: float = 1.0, group=None):
        self.session = session
        self.energy_scale = energy_scale
        self.group = group
        self.n_calls = 0
        self._buf: Optional[torch.Tensor] = None

    def __call__(self, atoms: AtomsX) -> Dict[str, torch.Tensor]:
        if atoms.cache is not None:
            return atoms.cache
        n = atoms.get_number_of_atoms()
        e, f_own, ids = self.session.energy_forces(atoms.positions, atoms.numbers)
        if self._buf is None or self._buf.shape[0] != 3 * n + 1:
            self._buf = torch.zeros(3 * n + 1, dtype=torch.float64, device=atoms.positions.device)
        buf = self._buf
        buf.zero_()
        buf[:3 * n].view(n, 3)[ids] = f_own.to(torch.float64)
        buf[3 * n] = e
        if self.session.world > 1:
            import torch.distributed as dist
            dist.all_reduce(buf, group=self.group)
        self.n_calls += 1
        out = {'energy': self.energy_scale * buf[3 * n].clone(), 'forces': self.energy_scale * buf[:3 * n].view(n, 3).clone()}
        atoms.cache = out
        return out


class VelocityVerletX:
    def __init__(self, timestep: float, calculator: CalculatorX):
        self.timestep, self.calculator = float(timestep), calculator

    @classmethod
    def create(cls, timestep, calculator):
        return cls(timestep, calculator)

    def step(self, atoms: AtomsX):
        masses = atoms.get_masses()[:, None]
        forces = self.calculator(atoms)['forces']
        p = atoms.get_momenta() + 0.5 * self.timestep * forces
        pos = atoms.get_positions()
        atoms_displaced = atoms.update_positions(pos + self.timestep * p / masses, update_neighbors=True)
        atoms_displaced = atoms_displaced.update_momenta(p)
        properties = self.calculator(atoms_displaced)
        forces = properties['forces']
        atoms_displaced = atoms_displaced.update_momenta(atoms_displaced.get_momenta() + 0.5 * self.timestep * forces)
        return self, atoms_displaced, properties


class NoseHooverX:
    """Nose-Hoover thermostat, mlff/mdx/integrator.py:17-83 (temperature in eV, ttime in units of the time step)."""

    def __init__(self, timestep: float, temperature: float, ttime: float, calculator: CalculatorX, zeta=0.0):
        self.timestep, self.temperature, self.ttime, self.calculator = float(timestep), float(temperature), float(ttime), calculator
        self.zeta = zeta

    @classmethod
    def create(cls, timestep, temperature, ttime, calculator):
        return cls(timestep, temperature, ttime, calculator)

    def target_Ekin(self, atoms):
        return 0.5 * (3.0 * atoms.get_number_of_atoms()) * self.temperature

    def Q(self, atoms):
        return 3.0 * atoms.get_number_of_atoms() * self.temperature * (self.ttime * self.timestep) ** 2

    def step(self, atoms: AtomsX):
        dt = self.timestep
        masses = atoms.get_masses()[:, None]
        accel = self.calculator(atoms)['forces'] / masses
        vel = atoms.get_velocities()
        x = atoms.get_positions() + vel * dt + (accel - self.zeta * vel) * (0.5 * dt ** 2)
        atoms_displaced = atoms.update_positions(x, update_neighbors=True)
        KE_0 = atoms_displaced.get_kinetic_energy()
        vel_half = vel + 0.5 * dt * (accel - self.zeta * vel)
        atoms_displaced = atoms_displaced.update_velocities(vel_half)
        properties = self.calculator(atoms_displaced)
        accel = properties['forces'] / masses
        zeta = self.zeta + 0.5 * dt * (1 / self.Q(atoms)) * (KE_0 - self.target_Ekin(atoms))
        zeta = zeta + 0.5 * dt * (1 / self.Q(atoms)) * (atoms_displaced.get_kinetic_energy() - self.target_Ekin(atoms))
        vel = (atoms_displaced.get_velocities() + 0.5 * dt * accel) / (1 + 0.5 * dt * zeta)
        atoms_displaced = atoms_displaced.update_velocities(vel)
        return NoseHooverX(self.timestep, self.temperature, self.ttime, self.calculator, zeta), atoms_displaced, properties


class LangevinX:
    """Langevin dynamics, mlff/mdx/integrator.py:86-159 (ASE's scheme: temperature in eV, friction in 1 / ASE time).
    The reference draws its noise from jax.random; here a torch generator on the atoms' device (the streams differ, the
    distributions do not)."""

    def __init__(self, timestep: float, temperature: float, friction: float, calculator: CalculatorX, fixcm: bool = True,
                 seed: int = 0):
        self.timestep, self.temperature, self.friction = float(timestep), float(temperature), float(friction)
        self.calculator, self.fixcm = calculator, fixcm
        self._gen: Optional[torch.Generator] = None
        self._seed = seed

    @classmethod
    def create(cls, timestep, temperature, friction, calculator, fixcm=True, seed=0):
        return cls(timestep, temperature, friction, calculator, fixcm, seed)

    def coefficients(self, atoms: AtomsX):
        dt, fr = self.timestep, self.friction
        sigma = torch.sqrt(2 * self.temperature * fr / atoms.get_masses())
        c1 = dt / 2. - dt * dt * fr / 8.
        c2 = dt * fr / 2 - dt * dt * fr * fr / 8.
        c3 = dt ** 0.5 * sigma / 2. - dt ** 1.5 * fr * sigma / 8.
        c5 = dt ** 1.5 * sigma / (2 * 3 ** 0.5)
        c4 = fr / 2. * c5
        return c1, c2, c3[:, None], c4[:, None], c5[:, None]

    def step(self, atoms: AtomsX):
        dev = atoms.positions.device
        if self._gen is None:
            self._gen = torch.Generator(device=dev).manual_seed(self._seed)
        masses = atoms.get_masses()[:, None]
        n = atoms.get_number_of_atoms()
        forces = self.calculator(atoms)['forces']
        vel = atoms.get_velocities()
        c1, c2, c3, c4, c5 = self.coefficients(atoms)
        xi = torch.randn(n, 3, generator=self._gen, dtype=forces.dtype, device=dev)
        eta = torch.randn(n, 3, generator=self._gen, dtype=forces.dtype, device=dev)
        rnd_pos = c5 * eta
        rnd_vel = c3 * xi - c4 * eta
        if self.fixcm:
            rnd_pos = rnd_pos - rnd_pos.sum(0, keepdim=True) / n
            rnd_vel = rnd_vel - (rnd_vel * masses).sum(0, keepdim=True) / (masses * n)
        vel = vel + (c1 * forces / masses - c2 * vel + rnd_vel)
        x = atoms.get_positions()
        atoms_displaced = atoms.update_positions(x + self.timestep * vel + rnd_pos)
        vel = (atoms_displaced.get_positions() - x - rnd_pos) / self.timestep
        properties = self.calculator(atoms_displaced)
        forces = properties['forces']
        vel = vel + (c1 * forces / masses - c2 * vel + rnd_vel)
        atoms_displaced = atoms_displaced.update_momenta(vel * masses)
        return self, atoms_displaced, properties


class BAOABLangevinX:
    """Langevin thermostat in the BAOAB splitting, mlff/mdx/integrator.py:162-233 (temperature in eV, gamma in 1 / ASE
    time): half kick, half drift, Ornstein-Uhlenbeck momentum update p <- c1 p + c2 sqrt(m) xi with c1 = exp(-gamma dt),
    c2 = sqrt(T (1 - c1^2)), optional removal of the centre-of-mass motion and of the rigid rotation (both WITHOUT
    re-scaling the temperature), half drift, half kick with the new forces.  As in the reference (`Normal.sample` draws a
    number per entry of the (n, 1) variance) ONE normal number per atom is shared by its three components."""

    def __init__(self, timestep: float, temperature: float, calculator: CalculatorX, gamma: float = 1e-1, fixcm: bool = True,
                 fixrot: bool = True, seed: int = 0):
        self.timestep, self.temperature, self.gamma = float(timestep), float(temperature), float(gamma)
        self.calculator, self.fixcm, self.fixrot = calculator, fixcm, fixrot
        self._gen: Optional[torch.Generator] = None
        self._seed = seed

    @classmethod
    def create(cls, timestep, temperature, calculator, gamma: float = 1e-1, fixcm: bool = True, fixrot: bool = True,
               seed: int = 0):
        return cls(timestep, temperature, calculator, gamma, fixcm, fixrot, seed)

    def _thermostat(self, atoms: AtomsX) -> AtomsX:
        dev = atoms.positions.device
        if self._gen is None:
            self._gen = torch.Generator(device=dev).manual_seed(self._seed)
        c1 = float(np.exp(-self.gamma * self.timestep))
        c2 = float(np.sqrt(self.temperature * (1.0 - c1 * c1)))
        masses = atoms.get_masses()[:, None]
        xi = torch.randn(atoms.get_number_of_atoms(), 1, generator=self._gen, dtype=masses.dtype, device=dev)
        atoms = atoms.update_momenta(c1 * atoms.get_momenta() + c2 * torch.sqrt(masses) * xi)
        if self.fixcm:
            atoms = zero_translation(atoms, preserve_temperature=False)
        if self.fixrot:
            atoms = zero_rotation(atoms, preserve_temperature=False)
        return atoms

    def step(self, atoms: AtomsX):
        half = 0.5 * self.timestep
        forces = self.calculator(atoms)['forces']
        atoms = atoms.update_momenta(atoms.get_momenta() + half * forces)
        atoms = atoms.update_positions(atoms.get_positions() + half * atoms.get_velocities())
        atoms = self._thermostat(atoms)
        atoms = atoms.update_positions(atoms.get_positions() + half * atoms.get_velocities())
        properties = self.calculator(atoms)
        atoms = atoms.update_momenta(atoms.get_momenta() + half * properties['forces'])
        return self, atoms, properties


class BerendsenX:
    """Berendsen velocity-rescaling thermostat, mlff/mdx/integrator.py:237-310 (temperature in eV, timestep and
    time_constant in ASE time units; the rescaling uses AtomsX.get_temperature(), i.e. 3n - 6 degrees of freedom)."""

    def __init__(self, timestep: float, temperature: float, calculator: CalculatorX, time_constant: float, fixcm: bool = True):
        self.timestep, self.temperature, self.calculator = float(timestep), float(temperature), calculator
        self.time_constant, self.fixcm = float(time_constant), fixcm

    @classmethod
    def create(cls, timestep, temperature, calculator, time_constant, fixcm: bool = True):
        return cls(timestep, temperature, calculator, time_constant, fixcm)

    def step(self, atoms: AtomsX):
        tr = self.timestep / self.time_constant
        scl = torch.sqrt(1.0 + tr * (self.temperature / atoms.get_temperature() - 1.0))
        atoms = atoms.update_momenta(scl * atoms.get_momenta())
        masses = atoms.get_masses()[:, None]
        forces = self.calculator(atoms)['forces']
        p = atoms.get_momenta() + 0.5 * self.timestep * forces
        if self.fixcm:
            p = p - p.sum(0) / float(len(p))
        atoms = atoms.update_positions(atoms.get_positions() + self.timestep * p / masses, update_neighbors=True)
        properties = self.calculator(atoms)
        atoms = atoms.update_momenta(p + 0.5 * self.timestep * properties['forces'])
        return self, atoms, properties


def scale_momenta(atoms: AtomsX, T0) -> AtomsX:
    """Uniform rescaling of the momenta to the temperature T0 (eV) -- mlff/mdx/utils.py:70-83."""
    return atoms.update_momenta(torch.sqrt(T0 / atoms.get_temperature()) * atoms.get_momenta())


def zero_translation(atoms: AtomsX, preserve_temperature: bool = True) -> AtomsX:
    """Remove the centre-of-mass motion: every atom loses the SAME velocity p_tot / m_tot -- mlff/mdx/utils.py:39-67."""
    T0 = atoms.get_temperature()
    atoms = atoms.update_velocities(atoms.get_velocities() - atoms.get_center_of_mass_velocity())
    return scale_momenta(atoms, T0) if preserve_temperature else atoms


def zero_rotation(atoms: AtomsX, preserve_temperature: bool = True) -> AtomsX:
    """Remove the rigid rotation that carries the total angular momentum: omega from L in the principal-axes frame
    (vanishing moments skipped), v <- v - omega x (r - r_com) -- mlff/mdx/utils.py:6-36."""
    T0 = atoms.get_temperature()
    Ip, basis = atoms.get_moments_of_inertia(vectors=True)
    Lp = basis @ atoms.get_angular_momentum()
    omega_p = torch.where(Ip > 0, Lp / torch.where(Ip > 0, Ip, torch.ones_like(Ip)), torch.zeros_like(Ip))
    omega = torch.linalg.solve(basis, omega_p)
    r = atoms.get_positions() - atoms.get_center_of_mass()
    atoms = atoms.update_velocities(atoms.get_velocities() - torch.linalg.cross(omega.expand_as(r), r))
    return scale_momenta(atoms, T0) if preserve_temperature else atoms


def simulate(integrator, atoms: AtomsX, steps: int):
    """Take `steps` steps; returns (atoms, {'kinetic_energy', 'potential_energy', 'temperature'} as (steps,) numpy arrays)."""
    ek, ep = [], []
    for _ in range(steps):
        integrator, atoms, prop = integrator.step(atoms)
        ek.append(atoms.get_kinetic_energy())
        ep.append(prop['energy'])
    ek_t, ep_t = torch.stack(ek).cpu().numpy(), torch.stack(ep).cpu().numpy()       # the only read-back of the run
    n = atoms.get_number_of_atoms()
    return atoms, {'kinetic_energy': ek_t, 'potential_energy': ep_t,
                   'temperature': 2.0 * ek_t / max(3 * n - 6, 1) / KB_EV}


class GradientDescent:
    """mlff/mdx/optimizer.py:16-52: x <- x - learning_rate * dE/dx until max_i |dE/dx_i| < tol."""

    def __init__(self, calculator: CalculatorX, learning_rate: float = 1e-2):
        self.calculator, self.learning_rate = calculator, float(learning_rate)

    @classmethod
    def create(cls, calculator, learning_rate: float = 1e-2):
        return cls(calculator, learning_rate)

    def step(self, atoms: AtomsX):
        grads = -self.calculator(atoms)['forces']
        return atoms.update_positions(atoms.get_positions() - self.learning_rate * grads), grads

    def minimize(self, atoms: AtomsX, max_steps: int = 1000, tol: float = 1e-2):
        for _ in range(max_steps):
            atoms, grads = self.step(atoms)
            if float(torch.linalg.norm(grads, dim=-1).max()) < tol:        # the reference's per-step read-back
                return atoms, grads
        raise RuntimeError('Optimization did not converge!')


class LBFGS:
    """Limited-memory BFGS relaxation with the settings the reference hands to jaxopt (mlff/mdx/optimizer.py:71-76,
    152-157): backtracking line search on the strong Wolfe conditions, initial inverse-Hessian scale gamma =
    s.y / y.y (`use_gamma`), step size capped at `max_stepsize` = 0.2, history 10, one iteration per `step`.
    jaxopt is not available in this image, so the sequence of iterates is this class's own (standard two-loop
    recursion); the contract is the reference's: `minimize` returns (relaxed atoms, forces) once the largest atomic
    force norm is below `tol`, raises RuntimeError otherwise.  Positions stay on the device; each iteration reads back
    a handful of scalars (as the reference's progress bar does)."""

    def __init__(self, calculator: CalculatorX, history_size: int = 10, max_stepsize: float = 0.2,
                 decrease_factor: float = 0.8, c1: float = 1e-4, c2: float = 0.9, max_linesearch: int = 15):
        self.calculator = calculator
        self.history_size, self.max_stepsize, self.decrease_factor = history_size, max_stepsize, decrease_factor
        self.c1, self.c2, self.max_linesearch = c1, c2, max_linesearch
        self.s_hist, self.y_hist = [], []
        self.n_iter = 0

    @classmethod
    def create(cls, atoms, calculator, save_dir: Optional[str] = None):
        opt = cls(calculator)
        opt.save_dir = save_dir
        return opt

    def _evaluate(self, atoms: AtomsX, x: torch.Tensor):
        a = atoms.update_positions(x)
        prop = self.calculator(a)
        return a, float(prop['energy']), -prop['forces']

    def _direction(self, g: torch.Tensor) -> torch.Tensor:
        q = g.clone()
        alphas = []
        for s_k, y_k in zip(reversed(self.s_hist), reversed(self.y_hist)):
            a_k = (s_k * q).sum() / (y_k * s_k).sum()
            q = q - a_k * y_k
            alphas.append(a_k)
        if self.s_hist:
            q = q * ((self.s_hist[-1] * self.y_hist[-1]).sum() / (self.y_hist[-1] * self.y_hist[-1]).sum())
        for (s_k, y_k), a_k in zip(zip(self.s_hist, self.y_hist), reversed(alphas)):
            b_k = (y_k * q).sum() / (y_k * s_k).sum()
            q = q + (a_k - b_k) * s_k
        return -q

    def step(self, atoms: AtomsX):
        """One L-BFGS iteration from `atoms`; returns (atoms at the new iterate, gradient there)."""
        x = atoms.get_positions()
        prop = self.calculator(atoms)          # cached when `atoms` comes out of the previous step
        e0, g0 = float(prop['energy']), -prop['forces']
        d = self._direction(g0)
        slope = float((d * g0).sum())
        if slope >= 0.0:                       # not a descent direction (stale curvature pairs): restart from steepest descent
            self.s_hist, self.y_hist = [], []
            d, slope = -g0, -float((g0 * g0).sum())
        # unit step of the quasi-Newton direction, capped so that no atom moves more than max_stepsize
        t = min(1.0, self.max_stepsize / max(float(torch.linalg.norm(d, dim=-1).max()), 1e-30))
        best = None
        for _ in range(self.max_linesearch):
            a1, e1, g1 = self._evaluate(atoms, x + t * d)
            armijo = e1 <= e0 + self.c1 * t * slope
            if armijo and best is None:
                best = (a1, g1, t)             # fall back to the first sufficient-decrease point
            if armijo and abs(float((d * g1).sum())) <= self.c2 * abs(slope):
                best = (a1, g1, t)
                break
            t *= self.decrease_factor
        if best is None:                       # no decrease found along d: tiny steepest-descent step, history dropped
            self.s_hist, self.y_hist = [], []
            a1, e1, g1 = self._evaluate(atoms, x - 1e-3 * self.max_stepsize * g0 / max(float(g0.abs().max()), 1e-30))
            best = (a1, g1, 0.0)
        a1, g1, _ = best
        s_k, y_k = a1.get_positions() - x, g1 - g0
        if float((s_k * y_k).sum()) > 1e-12:   # curvature condition keeps the inverse Hessian positive definite
            self.s_hist.append(s_k)
            self.y_hist.append(y_k)
            if len(self.s_hist) > self.history_size:
                self.s_hist.pop(0)
                self.y_hist.pop(0)
        self.n_iter += 1
        return a1, g1

    def minimize(self, atoms: AtomsX, max_steps: int, tol: float = 1e-3):
        max_force_norm = None
        for _ in range(max_steps):
            atoms, grads = self.step(atoms)
            max_force_norm = float(torch.linalg.norm(grads, dim=-1).max())
            if max_force_norm < tol:
                return atoms, -grads
        raise RuntimeError(f'BFGS optimization did not converge! {max_force_norm} (max. force norm ) > {tol} (tolerance).')


def maxwell_boltzmann_momenta(masses: torch.Tensor, temperature_K: float, seed: int = 0) -> torch.Tensor:
    """Momenta drawn from a Maxwell-Boltzmann distribution, centre-of-mass momentum removed (ASE convention)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    xi = torch.randn(masses.shape[0], 3, generator=g, dtype=torch.float64).to(masses.device)
    p = xi * torch.sqrt(masses * KB_EV * temperature_K)[:, None]
    return p - p.mean(0, keepdim=True)


def maxwell_boltzmann(masses: torch.Tensor, T0: float, seed: int = 0) -> torch.Tensor:
    """Momenta following the Maxwell-Boltzmann distribution at temperature T0 in eV: p = xi sqrt(m T0), xi ~ N(0, 1) --
    mlff/mdx/distributions.py:5-20 (there with a jax PRNG key; here a seeded torch generator -- the streams differ, the
    distribution does not; no centre-of-mass correction, as in the reference)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    xi = torch.randn(masses.shape[0], 3, generator=g, dtype=torch.float64).to(masses.device)
    return xi * torch.sqrt(masses.to(torch.float64) * T0)[:, None]
