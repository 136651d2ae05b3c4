"""MD under domain decomposition on the CPU: mlff_b200.dist.DDSession (brick selection, plan kept within a skin, rebuilt
beyond it) + mlff_b200.md.CalculatorDD (owned forces / energies summed over the ranks) + the velocity-Verlet loop, through
the emulated kernels -- one rank, and two ranks as real processes over gloo (world_size 2).  The GPU counterpart is
tests/test_gpu_parity.py::test_md_through_the_domain_decomposed_calculator; bench.py --gpus N --md-steps K runs the same
classes over NCCL."""
import os
import sys

import numpy as np
import torch

import so3k_oracle as o
from emu_engine import build_emu
from mlff_b200.capi import So3kLib, NL_ASE, STATUS_OVERFLOW
from mlff_b200 import dist as dd
from mlff_b200 import md
from mlff_b200 import params as P
from mlff_b200.config import So3kratesConfig

CFG = dict(F=32, n_layers=2, degrees=(1, 2), n_heads=4)
STEPS, SKIN, DT = 5, 0.06, 1.0 * md.FS


class EmuSessionEngine:
    """What DDSession needs of an engine (lib, h, device, neighbor_list, count_and_status), backed by the CPU-emulated
    kernels and torch CPU tensors.  TEST INFRASTRUCTURE."""

    def __init__(self, cfg, p):
        self.lib = So3kLib(build_emu())
        pc = So3kratesConfig(F=cfg.F, n_layer=cfg.n_layers, degrees=tuple(cfg.degrees), n_heads=cfg.n_heads)
        self.h = self.lib.create(pc.F, pc.n_rbf, pc.n_layer, pc.n_heads, list(pc.degrees), pc.r_cut)
        self.device = torch.device('cpu')
        self._blob = torch.from_numpy(P.pack_blob(pc, {k: np.asarray(v, np.float32) for k, v in p.items()}))
        self.lib.load_params(self.h, self._blob.data_ptr(), self._blob.numel())
        self._status = torch.zeros(1, dtype=torch.int32)

    def neighbor_list(self, positions, cutoff, capacity, cell=None, pbc=(False, False, False), z=None, mode=NL_ASE):
        pos = positions.to(torch.float64).contiguous()
        N = pos.shape[0]
        cell_h = np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(3, 3))
        pbc_h = np.ascontiguousarray(np.asarray(pbc, dtype=np.uint8).reshape(3))
        ii = torch.empty(capacity, dtype=torch.int32)
        jj = torch.empty(capacity, dtype=torch.int32)
        S = torch.empty(capacity, 3, dtype=torch.int32)
        count = torch.zeros(1, dtype=torch.int64)
        ws = torch.empty(self.lib.neighbor_list_workspace_bytes(N, capacity) + 1024, dtype=torch.uint8)
        self.lib.neighbor_list(mode, N, pos.data_ptr(), None, cell_h.ctypes.data, pbc_h.ctypes.data, cutoff, capacity,
                               ii.data_ptr(), jj.data_ptr(), S.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                               self._status.data_ptr(), 0)
        return ii, jj, S, count

    def count_and_status(self, count):
        s = int(self._status[0]) & ~STATUS_OVERFLOW
        self._status.zero_()
        assert s == 0, s
        return int(count[0])


def _system():
    sol = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'solid_0.npz'))
    cfg = o.OracleConfig(**CFG)
    p = o.init_params(cfg, 11, embed_scale=4.0)
    return sol['R'].astype(np.float64), sol['z'].astype(np.int32), sol['unit_cell'].astype(np.float64), cfg, p


def _run_md(world, rank, group=None):
    R, z, cell, cfg, p = _system()
    eng = EmuSessionEngine(cfg, p)
    sess = dd.DDSession(eng, cfg.n_layers, cell, world, rank, cfg.r_cut, skin=SKIN)
    masses = torch.as_tensor(np.where(z == 1, 1.008, 12.011 + 0.25 * z))
    mom = md.maxwell_boltzmann_momenta(masses, 600.0, seed=5)
    atoms = md.AtomsX(positions=torch.as_tensor(R), momenta=mom, masses=masses, numbers=torch.as_tensor(z), cell=cell,
                      pbc=(True, True, True)).init_domain_decomposition(sess)
    calc = md.CalculatorDD(sess, group=group)
    f0 = calc(atoms)['forces'].clone()
    e0 = float(calc(atoms)['energy'])
    atoms, st = md.simulate(md.VelocityVerletX.create(DT, calc), atoms, STEPS)
    return dict(f0=f0.numpy(), e0=e0, pos=atoms.positions.numpy().copy(), ep=st['potential_energy'],
                ek=st['kinetic_energy'], builds=sess.n_builds, calls=calc.n_calls)


_single = {}


def _single_rank():
    if not _single:
        _single.update(_run_md(1, 0))
    return _single


def test_single_rank_session_matches_the_oracle_and_conserves_energy():
    R, z, cell, cfg, p = _system()
    out = _single_rank()
    ii, jj, S = o.primitive_neighbor_list(R, cell, (True, True, True), cfg.r_cut)
    Eo, Fo, _ = o.energy_forces(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    Eo, Fo = float(Eo), Fo.numpy()
    assert abs(out['e0'] - Eo) <= 1e-5 * abs(Eo)
    assert np.abs(out['f0'] - Fo).max() < 1e-4
    # the plan is kept while no atom has moved more than skin / 2 and rebuilt beyond it
    assert 2 <= out['builds'] < out['calls']
    tot = out['ep'] + out['ek']
    assert np.abs(tot - tot[0]).max() <= 2e-3 * out['ek'].mean()


def _gloo_worker(rank, world, port, q):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.join(os.path.dirname(here), 'oracle'))
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        out = _run_md(world, rank)
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


def test_md_over_gloo_world_size_2_equals_the_single_rank_trajectory():
    """Two processes, each evaluating its brick of the same replicated AtomsX: the trajectory, the energies and the plan
    rebuild decisions equal those of the single-rank run."""
    import torch.multiprocessing as mp
    build_emu()
    ref = _single_rank()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    out = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert np.abs(out['f0'] - ref['f0']).max() < 2e-5
    assert np.abs(out['pos'] - ref['pos']).max() < 1e-5
    assert np.abs(out['ep'] - ref['ep']).max() <= 1e-5 * np.abs(ref['ep']).max()
    assert out['builds'] == ref['builds'] and out['calls'] == ref['calls']
