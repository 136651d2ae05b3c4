"""Domain decomposition ("mirror edges", mlff_b200/dist.py) on the CPU: the plan builder, the stepwise C ABI driven
through the emulated kernels, and the real multi-process exchange over gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from emu_engine import build_emu
from common import engine_flags
from mlff_b200.capi import So3kLib
from mlff_b200 import dist as dd
from mlff_b200 import params as P
from mlff_b200.config import So3kratesConfig


def _problem(solid):
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig()
    p = o.init_params(cfg, int(g['seed']), embed_scale=float(g['embed_scale']))
    return g, cfg, p


def _make_rank(lib, cfg, p, device):
    opts = {k: getattr(cfg, k) for k in ('layer_normalization', 'final_layer', 'residual_mlp_1', 'residual_mlp_2')}
    pc = So3kratesConfig(F=cfg.F, n_layer=cfg.n_layers, degrees=tuple(cfg.degrees), n_heads=cfg.n_heads, **opts)
    h = lib.create(pc.F, pc.n_rbf, pc.n_layer, pc.n_heads, list(pc.degrees), pc.r_cut, flags=engine_flags(cfg))
    blob = torch.from_numpy(P.pack_blob(pc, {k: np.asarray(v, np.float32) for k, v in p.items()})).to(device)
    lib.load_params(h, blob.data_ptr(), blob.numel())
    r = dd.DDRank(lib, h, pc.n_layer, device)
    r._blob = blob
    return r


def test_plan_partitions_the_edge_list(solid):
    g, cfg, p = _problem(solid)
    ii, jj, S = (torch.as_tensor(g[k]) for k in ('idx_i', 'idx_j', 'shifts'))
    pos = torch.as_tensor(solid['R'])
    for world in (2, 3):
        owner = dd.assign_owners_slab(pos, solid['unit_cell'], world, axis=2)
        plans = [dd.build_plan(ii, jj, S, owner, r, world) for r in range(world)]
        assert sum(pl.n_own for pl in plans) == 110
        # owned rows cover every edge exactly once; mirrors are exactly the cross-rank edges seen from the sender
        n_owned_edges = sum(int((pl.idx_i < pl.n_own).sum()) for pl in plans)
        assert n_owned_edges == len(ii)
        cross = int((owner[ii] != owner[jj]).sum())
        assert sum(int((pl.idx_i >= pl.n_own).sum()) for pl in plans) == cross
        for pl in plans:
            assert sum(pl.send_counts) == len(pl.send_idx) and sum(pl.recv_counts) == pl.n_ghost
            assert pl.send_counts[pl.rank] == 0 and pl.recv_counts[pl.rank] == 0
            assert (pl.send_idx < pl.n_own).all()
            for q in plans:                       # what I receive from q is what q sends to me, in the same order
                s0 = sum(q.send_counts[:pl.rank])
                sent = q.local_atoms[q.send_idx[s0: s0 + q.send_counts[pl.rank]]]
                r0 = pl.n_own + sum(pl.recv_counts[:q.rank])
                assert torch.equal(sent, pl.local_atoms[r0: r0 + pl.recv_counts[q.rank]])


@pytest.mark.parametrize('world', [2, 3])
def test_domain_decomposed_energy_forces_match_oracle(solid, world):
    """Virtual ranks in one process, emulated kernels: E and F of the decomposed evaluation == single-domain oracle."""
    g, cfg, p = _problem(solid)
    lib = So3kLib(build_emu())
    dev = torch.device('cpu')
    pos = torch.as_tensor(solid['R'])
    z = torch.as_tensor(solid['z'])
    ii, jj, S = (torch.as_tensor(g[k]) for k in ('idx_i', 'idx_j', 'shifts'))
    owner = dd.assign_owners_slab(pos, solid['unit_cell'], world, axis=2)
    ranks = []
    for r in range(world):
        rk = _make_rank(lib, cfg, p, dev)
        rk.begin(dd.build_plan(ii, jj, S, owner, r, world), pos, z, solid['unit_cell'])
        ranks.append(rk)
    dd.run_energy_forces(ranks, dd.exchange_inprocess)
    E = sum(float(rk.e_atom[:rk.plan.n_own].double().sum()) for rk in ranks)
    F = np.zeros((110, 3), np.float32)
    for rk in ranks:
        assert int(rk.status[0]) == 0
        F[rk.plan.local_atoms[:rk.plan.n_own].numpy()] = rk.forces[:rk.plan.n_own].numpy()
    assert abs(E - float(g['E'])) <= 1e-5 * abs(float(g['E']))
    assert np.abs(F - g['F']).max() < 1e-4


def test_domain_decomposition_with_optional_layer_branches(solid):
    """LayerNorm / ResidualMLP are node-local, so the ghost exchanges of the default model carry them unchanged."""
    g = golden('oracle_solid_pbc.npz')
    cfg = o.OracleConfig(F=36, n_layers=2, layer_normalization=True, final_layer=True, residual_mlp_1=True,
                         residual_mlp_2=True)
    p = o.init_params(cfg, 3, embed_scale=4.0)
    lib = So3kLib(build_emu())
    dev = torch.device('cpu')
    pos = torch.as_tensor(solid['R'])
    z = torch.as_tensor(solid['z'])
    ii, jj, S = (torch.as_tensor(g[k]) for k in ('idx_i', 'idx_j', 'shifts'))
    Eo, Fo, _ = o.energy_forces(cfg, p, solid['R'], solid['z'], g['idx_i'], g['idx_j'], cell=solid['unit_cell'],
                             cell_offset=g['shifts'])
    owner = dd.assign_owners_slab(pos, solid['unit_cell'], 2, axis=2)
    ranks = []
    for r in range(2):
        rk = _make_rank(lib, cfg, p, dev)
        rk.begin(dd.build_plan(ii, jj, S, owner, r, 2), pos, z, solid['unit_cell'])
        ranks.append(rk)
    dd.run_energy_forces(ranks, dd.exchange_inprocess)
    E = sum(float(rk.e_atom[:rk.plan.n_own].double().sum()) for rk in ranks)
    F = np.zeros((110, 3), np.float32)
    for rk in ranks:
        assert int(rk.status[0]) == 0
        F[rk.plan.local_atoms[:rk.plan.n_own].numpy()] = rk.forces[:rk.plan.n_own].numpy()
    assert abs(E - float(Eo)) <= 1e-5 * max(abs(float(Eo)), 1.0)
    assert np.abs(F - Fo.numpy()).max() < 1e-4


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sol = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'solid_0.npz'))
        solid = {'R': sol['R'], 'z': sol['z'], 'unit_cell': sol['unit_cell']}
        g, cfg, p = _problem(solid)
        lib = So3kLib(build_emu())
        pos, z = torch.as_tensor(solid['R']), torch.as_tensor(solid['z'])
        ii, jj, S = (torch.as_tensor(g[k]) for k in ('idx_i', 'idx_j', 'shifts'))
        owner = dd.assign_owners_slab(pos, solid['unit_cell'], world, axis=2)
        rk = _make_rank(lib, cfg, p, torch.device('cpu'))
        rk.begin(dd.build_plan(ii, jj, S, owner, rank, world), pos, z, solid['unit_cell'])
        dd.run_energy_forces([rk], dd.exchange_distributed)
        e = rk.e_atom[:rk.plan.n_own].double().sum().reshape(1)
        dist.all_reduce(e)
        F = torch.zeros(110, 3)
        F[rk.plan.local_atoms[:rk.plan.n_own]] = rk.forces[:rk.plan.n_own]
        dist.all_reduce(F)
        if rank == 0:
            q.put((float(e), F.numpy(), float(g['E']), g['F']))
    finally:
        dist.destroy_process_group()


def test_domain_decomposition_over_gloo_world_size_2():
    """The N > 1 path end to end on the CPU: one process per rank, halo rows exchanged with torch.distributed (gloo)."""
    import torch.multiprocessing as mp
    build_emu()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    E, F, Eref, Fref = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert abs(E - Eref) <= 1e-5 * abs(Eref)
    assert np.abs(F - Fref).max() < 1e-4


def test_halo_subset_gives_the_same_plan_edges(solid):
    """Building a rank's neighbour list only on the atoms of its brick + r_cut yields the same owned rows and mirror
    edges (as global (i, j, S) sets) as building it on the whole system."""
    pos_np, cell = solid['R'], solid['unit_cell']
    pos = torch.as_tensor(pos_np)
    world = 3
    oi, oj, oS = o.primitive_neighbor_list(pos_np, cell, (1, 1, 1), 5.0)
    owner = dd.assign_owners_bricks(pos, cell, world)
    assert dd.brick_grid(cell, world) == (1, 1, 3)
    for rank in range(world):
        full = dd.build_plan(torch.as_tensor(oi), torch.as_tensor(oj), torch.as_tensor(oS), owner, rank, world)
        keep = dd.brick_halo_mask(pos, cell, world, rank, 5.0)
        sub = torch.nonzero(keep).reshape(-1)
        si, sj, sS = o.primitive_neighbor_list(pos_np[sub.numpy()], cell, (1, 1, 1), 5.0)
        part = dd.build_plan(torch.as_tensor(si), torch.as_tensor(sj), torch.as_tensor(sS), owner[sub], rank, world)

        def edge_set(pl, ids):
            gi, gj = ids[pl.local_atoms[pl.idx_i]], ids[pl.local_atoms[pl.idx_j]]
            return sorted(zip(gi.tolist(), gj.tolist(), map(tuple, pl.shifts.tolist())))
        assert edge_set(full, torch.arange(110)) == edge_set(part, sub)
        assert torch.equal(sub[part.local_atoms], full.local_atoms)
        assert part.send_counts == full.send_counts and part.recv_counts == full.recv_counts


def test_brick_select_equals_mask_and_owners():
    rng = np.random.default_rng(3)
    cell = np.array([[31.0, 0.0, 0.0], [3.0, 29.0, 0.0], [-2.0, 4.0, 33.0]])
    pos = torch.as_tensor(rng.uniform(-40, 80, size=(5000, 3)))
    for world in (2, 3, 4, 8):
        for rank in range(world):
            sub, own = dd.brick_select(pos, cell, world, rank, 5.0)
            keep = dd.brick_halo_mask(pos, cell, world, rank, 5.0)
            assert torch.equal(sub, torch.nonzero(keep).reshape(-1))
            assert torch.equal(own, dd.assign_owners_bricks(pos[sub], cell, world))
