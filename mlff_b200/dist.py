"""Spatial domain decomposition of ONE large periodic system over several GPUs (the
reference has no multi-device path).

Design ("mirror edges"): the atoms are partitioned into slabs; a rank's local problem is its OWNED atoms followed by
GHOST atoms (every sender of an edge received by an owned atom), and its local edge list is all edges received by
owned atoms PLUS their reverses (receiver ghost, sender owned).  The local list is therefore symmetric, so every
reduction the forward and the analytic backward need on an owned atom is local -- the kernels run unchanged -- and
the only communication is one owner->ghost row exchange per layer in each direction of the sweep
(forward: x_{t+1}, chi_{t+1}; backward: xbar1, chibar1), plus a scalar all-reduce of the energy.  Nothing is ever
accumulated from ghosts back into owners.

A rank prepares its problem from the atoms of its brick + r_cut only (brick_select): subset neighbour list on the
device cell list, plan (owned rows, mirror edges, ghost order, send lists) from it -- O(local), and without
communication, because both sides of every exchange derive the same (needing rank, atom id)-sorted lists from
bit-identical distance tests.  DDSession keeps a plan across steps while no atom has moved more than skin / 2 (the
list is then built with r_cut + skin; the extra pairs contribute exactly zero through the cosine cutoff), which takes
the preparation -- a quarter of an 8-GPU step -- out of most steps.  All tensor plumbing here is torch (device memory,
NCCL); the computation is libso3k's so3k_dd_* entry points.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .capi import So3kLib


def brick_grid(cell: np.ndarray, world: int) -> tuple:
    """(px, py, pz) with px*py*pz == world minimising the brick surface for this cell (e.g. 8 -> 2x2x2 in a cube)."""
    L = np.linalg.norm(np.asarray(cell, dtype=np.float64).reshape(3, 3), axis=1)
    best, best_cost = (world, 1, 1), None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // (px * py)
            a, b, c = L[0] / px, L[1] / py, L[2] / pz
            cost = (px > 1) * b * c + (py > 1) * a * c + (pz > 1) * a * b      # exposed face area per brick
            if best_cost is None or cost < best_cost - 1e-12:
                best, best_cost = (px, py, pz), cost
    return best


def _frac(pos: torch.Tensor, inv_np: np.ndarray) -> torch.Tensor:
    """Fractional coordinates wrapped into [0, 1), evaluated element-wise (no GEMM): every rank must get bit-identical
    values for an atom whatever the size of the array it sits in, or two ranks could disagree about its owner."""
    inv = torch.as_tensor(inv_np, dtype=pos.dtype, device=pos.device)
    frac = pos[:, 0:1] * inv[0] + pos[:, 1:2] * inv[1] + pos[:, 2:3] * inv[2]
    return frac - torch.floor(frac)


def assign_owners_bricks(pos: torch.Tensor, cell: np.ndarray, world: int) -> torch.Tensor:
    """owner[a] = index of the brick (fractional-coordinate grid of brick_grid(cell, world)) that contains atom a."""
    px, py, pz = brick_grid(cell, world)
    frac = _frac(pos, np.linalg.inv(np.asarray(cell, dtype=np.float64)))
    cx = torch.clamp((frac[:, 0] * px).to(torch.int64), 0, px - 1)
    cy = torch.clamp((frac[:, 1] * py).to(torch.int64), 0, py - 1)
    cz = torch.clamp((frac[:, 2] * pz).to(torch.int64), 0, pz - 1)
    return (cx * py + cy) * pz + cz


def brick_halo_mask(pos: torch.Tensor, cell: np.ndarray, world: int, rank: int, margin: float) -> torch.Tensor:
    """Atoms inside brick `rank` expanded by `margin` (a length, e.g. r_cut) on every side, with periodic wrap.
    These are the only atoms the rank needs to build its rows and its mirror edges (senders of owned rows and
    receivers of mirror edges are within r_cut of an owned atom)."""
    px, py, pz = brick_grid(cell, world)
    c = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    inv_np = np.linalg.inv(c)
    frac = _frac(pos, inv_np)
    heights = 1.0 / np.linalg.norm(inv_np, axis=0)                 # face-to-face distances
    idx = (rank // (py * pz), (rank // pz) % py, rank % pz)
    keep = torch.ones(pos.shape[0], dtype=torch.bool, device=pos.device)
    for a, (n, k) in enumerate(zip((px, py, pz), idx)):
        if n == 1:
            continue
        lo, hi = k / n, (k + 1) / n
        m = margin / heights[a] * (1.0 + 1e-9) + 1e-12
        centre, half = 0.5 * (lo + hi), 0.5 * (hi - lo) + m
        if half >= 0.5:
            continue
        dist = torch.abs(frac[:, a] - centre)
        dist = torch.minimum(dist, 1.0 - dist)                      # periodic distance to the brick centre
        keep &= dist <= half
    return keep


def brick_select(pos: torch.Tensor, cell: np.ndarray, world: int, rank: int, margin: float):
    """(sub, owner_sub): ascending global ids of the atoms inside brick `rank` + margin (brick_halo_mask) and the owning
    brick of each of them (assign_owners_bricks) from ONE evaluation of the fractional coordinates -- the per-step
    preparation of a rank is launch-bound, so the number of small device ops matters."""
    px, py, pz = brick_grid(cell, world)
    c = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    inv_np = np.linalg.inv(c)
    frac = _frac(pos, inv_np)
    heights = 1.0 / np.linalg.norm(inv_np, axis=0)
    grid = torch.tensor([px, py, pz], dtype=pos.dtype, device=pos.device)
    cells = torch.minimum((frac * grid).to(torch.int64), (grid - 1).to(torch.int64)).clamp_min_(0)
    owner = (cells[:, 0] * py + cells[:, 1]) * pz + cells[:, 2]
    idx = (rank // (py * pz), (rank // pz) % py, rank % pz)
    centre = np.zeros(3)
    half = np.full(3, 2.0)                                          # > 0.5: no restriction along this axis
    for a_, (n, k) in enumerate(zip((px, py, pz), idx)):
        if n > 1:
            m = margin / heights[a_] * (1.0 + 1e-9) + 1e-12
            centre[a_], half[a_] = (k + 0.5) / n, 0.5 / n + m
    d = torch.abs(frac - torch.as_tensor(centre, dtype=pos.dtype, device=pos.device))
    d = torch.minimum(d, 1.0 - d)                                   # periodic distance to the brick centre
    keep = (d <= torch.as_tensor(half, dtype=pos.dtype, device=pos.device)).all(dim=1)
    sub = torch.nonzero(keep).reshape(-1)
    return sub, owner[sub]


def assign_owners_slab(pos: torch.Tensor, cell: np.ndarray, world: int, axis: int = 0) -> torch.Tensor:
    """owner[a] = slab index of atom a along `axis` (fractional coordinate wrapped into [0,1))."""
    frac = _frac(pos, np.linalg.inv(np.asarray(cell, dtype=np.float64)))[:, axis]
    return torch.clamp((frac * world).to(torch.int64), 0, world - 1)


@dataclass
class DDPlan:
    rank: int
    world: int
    n_own: int
    n_ghost: int
    local_atoms: torch.Tensor          # (n_own + n_ghost) global ids: owned (ascending), then ghosts by (owner, id)
    idx_i: torch.Tensor                # (P_loc) local receiver
    idx_j: torch.Tensor                # (P_loc) local sender
    shifts: torch.Tensor               # (P_loc, 3)
    send_idx: torch.Tensor             # local (owned) rows to send, grouped by destination rank
    send_counts: List[int]             # per destination rank
    recv_counts: List[int]             # per source rank (== layout of the ghost block)

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_ghost


def build_plan(ii: torch.Tensor, jj: torch.Tensor, S: torch.Tensor, owner: torch.Tensor, rank: int, world: int,
               check: bool = True) -> DDPlan:
    """ii, jj (P) int64 global ids of the valid edges (receiver, sender), S (P,3), owner (N).
    check=False skips the symmetry assertion (one host synchronisation less per step; an asymmetric list is still
    reported by the kernels through SO3K_STATUS_ASYMMETRIC)."""
    dev = ii.device
    N = owner.shape[0]
    oi, oj = owner[ii], owner[jj]
    own_recv = oi == rank
    mirror = (oj == rank) & (oi != rank)                     # receiver is someone else's atom, sender is mine
    # ghosts: senders of my rows that I do not own, ordered by (owner, global id)
    gflag = torch.zeros(N, dtype=torch.bool, device=dev)     # (flag + nonzero: cheaper than a sort-based unique)
    gflag[jj[own_recv & (oj != rank)]] = True
    gh = torch.nonzero(gflag).reshape(-1)                    # ascending global id
    gh = gh[torch.sort(owner[gh], stable=True).indices]      # -> ordered by (owner, global id)
    owned = torch.nonzero(owner == rank).reshape(-1)
    local_atoms = torch.cat([owned, gh])
    g2l = torch.full((N,), -1, dtype=torch.int64, device=dev)
    g2l[local_atoms] = torch.arange(local_atoms.shape[0], device=dev)
    sel = own_recv | mirror
    li, lj = g2l[ii[sel]], g2l[jj[sel]]
    if check:
        assert bool(((li >= 0) & (lj >= 0)).all()), 'mirror edge whose receiver is not a ghost (asymmetric list?)'
    else:
        li, lj = li.clamp_min(0), lj.clamp_min(0)
    counts = torch.stack([torch.bincount(owner[gh], minlength=world)[:world],
                          torch.zeros(world, dtype=torch.int64, device=dev)])
    # what the others need from me: unique (needing rank, my atom) over edges received elsewhere and sent by me
    nflag = torch.zeros(world * N, dtype=torch.bool, device=dev)
    nflag[oi[mirror] * N + jj[mirror]] = True
    need = torch.nonzero(nflag).reshape(-1)                  # sorted by (needing rank, global id)
    need_rank, need_atom = need // N, need % N
    counts[1] = torch.bincount(need_rank, minlength=world)[:world]
    recv_counts, send_counts = counts.tolist()                # one host synchronisation for both
    send_idx = g2l[need_atom]
    return DDPlan(rank, world, int(owned.shape[0]), int(gh.shape[0]), local_atoms, li, lj, S[sel], send_idx,
                  send_counts, recv_counts)


class DDRank:
    """State of one rank of a domain-decomposed evaluation (buffers + the stepwise C-ABI calls)."""

    def __init__(self, lib: So3kLib, handle: int, n_layers: int, device: torch.device):
        self.lib, self.h, self.L, self.device = lib, handle, n_layers, device
        self.ws: Optional[torch.Tensor] = None
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.plan: Optional[DDPlan] = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0

    def begin(self, plan: DDPlan, pos: torch.Tensor, z: torch.Tensor, cell: np.ndarray):
        self.plan = plan
        dev = self.device
        la = plan.local_atoms
        self.R = pos[la].to(torch.float32).contiguous()
        self.z = z[la].to(torch.int32).contiguous()
        self.ii = plan.idx_i.to(torch.int32).contiguous()
        self.jj = plan.idx_j.to(torch.int32).contiguous()
        self.S = plan.shifts.to(torch.int32).contiguous()
        self.cell = torch.as_tensor(np.asarray(cell, dtype=np.float32).reshape(1, 3, 3), device=dev).contiguous()
        self.N, self.P = plan.n_local, int(self.ii.shape[0])
        nbytes = self.lib.workspace_bytes(self.h, self.N, self.P)
        if self.ws is None or self.ws.numel() < nbytes:
            self.ws = torch.zeros(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=dev)
        self.e_atom = torch.zeros(self.N, dtype=torch.float32, device=dev)
        self.forces = torch.zeros(self.N, 3, dtype=torch.float32, device=dev)
        self.lib.dd_begin(self.h, self.N, self.P, self.R.data_ptr(), self.z.data_ptr(), self.ii.data_ptr(),
                          self.jj.data_ptr(), self.cell.data_ptr(), self.S.data_ptr(), self.ws.data_ptr(),
                          self.ws.numel(), self.status.data_ptr(), self._stream())

    def begin_local(self, plan: DDPlan, R_local: torch.Tensor, z_local: torch.Tensor, cell: np.ndarray):
        """`begin` for callers that keep the plan across steps: R_local / z_local are already in the plan's local order;
        the int32 index arrays are converted once per plan."""
        if getattr(self, '_plan_arrays_of', None) is not plan:
            self.ii = plan.idx_i.to(torch.int32).contiguous()
            self.jj = plan.idx_j.to(torch.int32).contiguous()
            self.S = plan.shifts.to(torch.int32).contiguous()
            self.cell = torch.as_tensor(np.asarray(cell, dtype=np.float32).reshape(1, 3, 3), device=self.device).contiguous()
            self._plan_arrays_of = plan
        self.plan = plan
        dev = self.device
        self.R = R_local.to(torch.float32).contiguous()
        self.z = z_local.to(torch.int32).contiguous()
        self.N, self.P = plan.n_local, int(self.ii.shape[0])
        nbytes = self.lib.workspace_bytes(self.h, self.N, self.P)
        if self.ws is None or self.ws.numel() < nbytes:
            self.ws = torch.zeros(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=dev)
        if getattr(self, 'e_atom', None) is None or self.e_atom.shape[0] != self.N:
            self.e_atom = torch.zeros(self.N, dtype=torch.float32, device=dev)
            self.forces = torch.zeros(self.N, 3, dtype=torch.float32, device=dev)
        self.lib.dd_begin(self.h, self.N, self.P, self.R.data_ptr(), self.z.data_ptr(), self.ii.data_ptr(),
                          self.jj.data_ptr(), self.cell.data_ptr(), self.S.data_ptr(), self.ws.data_ptr(),
                          self.ws.numel(), self.status.data_ptr(), self._stream())

    def stage(self, stage: int, t: int = 0):
        out = {1: self.e_atom, 4: self.forces}.get(stage)
        # node-level work on the owned rows only (they come first): a ghost's outputs are its owner's
        self.lib.dd_stage_owned(self.h, stage, t, self.N, self.plan.n_own, self.P, self.z.data_ptr(),
                                None if out is None else out.data_ptr(), self.ws.data_ptr(), self.status.data_ptr(),
                                self._stream())

    def buffer(self, which: int, t: int) -> torch.Tensor:
        """(N_local, width) float32 view of x_t / chi_t / xbar1 / chibar1 inside the workspace."""
        off, rf = self.lib.dd_buffer(self.h, which, t, self.N, self.P)
        return self.ws[off: off + 4 * self.N * rf].view(torch.float32).view(self.N, rf)


Exchange = Callable[[Sequence[DDRank], Sequence[int], int], None]


def exchange_inprocess(ranks: Sequence[DDRank], which: Sequence[int], t: int) -> None:
    """All ranks live in this process (tests): copy owner rows into the ghost blocks directly."""
    for w in which:
        bufs = [r.buffer(w, t) for r in ranks]
        for r, dst in zip(ranks, bufs):
            p = r.plan
            off = p.n_own
            for src_rank in range(p.world):
                cnt = p.recv_counts[src_rank]
                if cnt == 0:
                    continue
                sp = ranks[src_rank].plan
                s0 = sum(sp.send_counts[:p.rank])
                rows = sp.send_idx[s0: s0 + sp.send_counts[p.rank]]
                dst[off: off + cnt] = bufs[src_rank][rows].to(dst.device)
                off += cnt


def exchange_distributed(ranks: Sequence[DDRank], which: Sequence[int], t: int) -> None:
    """One rank per process: the owned rows the peers need of both buffers of the exchange go into one message per peer
    (so3k_dd_pack), are exchanged (NCCL all_to_all_single; point-to-point sends on backends without all-to-all, e.g. gloo in
    the CPU tests) and land in the ghost blocks (so3k_dd_unpack)."""
    import torch.distributed as dist
    (r,) = ranks
    p = r.plan
    pair = {(0, 1): 0, (2, 3): 1}[tuple(which)]
    if getattr(r, '_halo_of', None) is not p:
        # per plan: int32 send list on the device and the two message buffers (one row = [x row | chi row] of an atom)
        bufs = [r.buffer(w, t) for w in which]
        width = sum(b.shape[1] for b in bufs)
        dev = bufs[0].device
        r._send_idx32 = p.send_idx.to(dev, torch.int32).contiguous()
        r._send = torch.empty(int(r._send_idx32.shape[0]), width, dtype=torch.float32, device=dev)
        r._recv = torch.empty(p.n_ghost, width, dtype=torch.float32, device=dev)
        r._halo_of = p
    send, recv = r._send, r._recv
    # gather the owned rows the peers need of both buffers into one message (one kernel), exchange, scatter the received
    # rows into the ghost blocks (one kernel)
    r.lib.dd_pack(r.h, pair, t, r.N, r.P, r._send_idx32.data_ptr(), int(send.shape[0]), send.data_ptr(), r.ws.data_ptr(),
                  r._stream())
    if dist.get_backend() == 'nccl':
        dist.all_to_all_single(recv, send, output_split_sizes=p.recv_counts, input_split_sizes=p.send_counts)
    else:
        reqs, so, ro = [], 0, 0
        for peer in range(p.world):
            sc, rc = p.send_counts[peer], p.recv_counts[peer]
            if peer != p.rank:
                if sc:
                    reqs.append(dist.isend(send[so: so + sc].contiguous(), peer))
                if rc:
                    reqs.append(dist.irecv(recv[ro: ro + rc], peer))
            so += sc
            ro += rc
        for q in reqs:
            q.wait()
    r.lib.dd_unpack(r.h, pair, t, r.N, p.n_own, r.P, recv.data_ptr(), r.ws.data_ptr(), r._stream())


def run_energy_forces(ranks: Sequence[DDRank], exchange: Exchange) -> None:
    """Lock-step sweep over the stages; afterwards rank.e_atom[:n_own] and rank.forces[:n_own] are final."""
    L = ranks[0].L
    for t in range(L):
        for r in ranks:
            r.stage(0, t)
        if t + 1 < L:                             # (x_L, chi_L of a ghost are never read: its read-out and its
            exchange(ranks, (0, 1), t + 1)        #  cotangents belong to the owner)  x_{t+1}, chi_{t+1} of the ghosts
    for r in ranks:
        r.stage(1)
    for t in range(L - 1, -1, -1):
        for r in ranks:
            r.stage(2, t)
        exchange(ranks, (2, 3), t)                # xbar1, chibar1 of the ghosts
        for r in ranks:
            r.stage(3, t)
    for r in ranks:
        r.stage(4)


class DDSession:
    """One rank's domain-decomposed energy + force evaluation of a periodic system, with the plan (brick subset,
    neighbour list, owned / ghost order, mirror edges, send lists) RE-USED across calls while no atom has moved more than
    skin / 2 since it was built -- the skin rule of the reference's MD lists (glp quadratic_neighbor_list,
    mlff/mdx/atoms.py:365-383, 488-510).  skin = 0 rebuilds on every call."""

    def __init__(self, engine, n_layers: int, cell: np.ndarray, world: int, rank: int, cutoff: float, skin: float = 0.0,
                 cuda_graph: bool = False,
                 capacity: int = 0):
        from .capi import NL_ASE
        self.engine, self.cell, self.world, self.rank = engine, np.asarray(cell, dtype=np.float64), world, rank
        self.cutoff, self.skin, self.mode = float(cutoff), float(skin), NL_ASE
        self.dd = DDRank(engine.lib, engine.h, n_layers, engine.device)
        self.capacity = int(capacity)
        self.plan: Optional[DDPlan] = None
        self.ref: Optional[torch.Tensor] = None
        self.n_builds = 0
        self.marks: Optional[list] = None            # (label, event) of the last call when timing is requested
        # cuda_graph: while a plan is kept, the whole sweep (local graph + embedding, layer stages, halo pack / NCCL
        # exchange / unpack, energy sum: ~400 short launches at 8 ranks) is captured once and replayed as ONE CUDA graph
        # -- libso3k only enqueues and never allocates, NCCL collectives are capturable, and every rank takes the same
        # rebuild decisions (they all see the whole position array), so the ranks capture and replay in lock step
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        self._graph_plan = None

    def _mark(self, label: str):
        if self.marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((label, ev))

    def _rebuild(self, pos: torch.Tensor):
        eng = self.engine
        rc = self.cutoff + self.skin
        sub, owner = brick_select(pos, self.cell, self.world, self.rank, rc)      # ascending: subset order == global order
        pos_s = pos[sub]
        if self.capacity <= 0:
            self.capacity = 64 * int(sub.shape[0]) + 4096
        while True:
            ii, jj, S, count = eng.neighbor_list(pos_s, rc, self.capacity, cell=self.cell, pbc=(True, True, True),
                                                 mode=self.mode)
            nv = eng.count_and_status(count)                    # plan sizes are host values (one read per rebuild)
            if nv <= self.capacity:
                break
            self.capacity = int(1.25 * nv) + 4096
        self._mark('subset + neighbour list')
        self.plan = build_plan(ii[:nv].long(), jj[:nv].long(), S[:nv], owner, self.rank, self.world, check=False)
        self.atoms = sub[self.plan.local_atoms]                 # global ids of the local rows (owned, then ghosts)
        self.ref = pos.clone()
        self.n_builds += 1
        self._mark('plan')

    def energy_forces(self, pos: torch.Tensor, z: torch.Tensor, timing: bool = False):
        """pos (N,3) float64 and z (N) int32: the WHOLE system, replicated on every rank (device tensors).  Returns the
        sum of this rank's owned per-atom energies (float64 device scalar, to be all-reduced by the caller), the forces
        of the owned atoms (n_own, 3) and their global ids."""
        self.marks = [] if timing else None
        self._mark('start')
        rebuild = self.plan is None or self.skin <= 0.0
        if not rebuild:
            d2 = ((pos - self.ref) ** 2).sum(-1).max()
            rebuild = float(d2.item()) > (0.5 * self.skin) ** 2
        if rebuild:
            self._rebuild(pos)
        plan = self.plan
        exchange = exchange_distributed if self.world > 1 else exchange_inprocess
        if self.cuda_graph and self.skin > 0.0 and not timing:
            if self._graph_plan is not plan:
                self._capture(plan, pos, z, exchange)
            torch.index_select(pos, 0, self.atoms, out=self._pos_l)
            self._R_l.copy_(self._pos_l)
            torch.index_select(z, 0, self.atoms, out=self._z_l)
            self._graph.replay()
            return self._e_l, self.dd.forces[:plan.n_own], self.atoms[:plan.n_own]
        # local arrays straight from the global ones (the plan's local order), graph + embedding, then the layer sweep
        self.dd.begin_local(plan, pos[self.atoms], z[self.atoms], self.cell)
        self._mark('begin (local arrays, CSR, embedding)')
        run_energy_forces([self.dd], exchange)
        self._mark('layers + exchanges')
        e = self.dd.e_atom[:plan.n_own].double().sum()
        return e, self.dd.forces[:plan.n_own], self.atoms[:plan.n_own]

    def _capture(self, plan, pos, z, exchange):
        """One eager sweep on a side stream (allocates the workspace, warms NCCL up for these message sizes), then the
        capture of the same calls on static local buffers."""
        dev = self.engine.device
        self._graph = None                                          # release the previous plan's graph and its pool
        self._pos_l = pos[self.atoms].contiguous()
        self._R_l = self._pos_l.to(torch.float32)
        self._z_l = z[self.atoms].to(torch.int32).contiguous()
        if z.dtype != torch.int32:
            raise TypeError('cuda_graph=True needs int32 atomic numbers (they are gathered into a static buffer)')

        def sweep():
            self.dd.begin_local(plan, self._R_l, self._z_l, self.cell)
            run_energy_forces([self.dd], exchange)
            return self.dd.e_atom[:plan.n_own].double().sum()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            sweep()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = self.engine.lib.launch_count()
        with torch.cuda.graph(g):
            self._e_l = sweep()
        self.captured_launches = self.engine.lib.launch_count() - n0          # libso3k kernels one replay runs
        self._graph, self._graph_plan = g, plan

    def close(self):
        """Release the captured graph (it holds NCCL kernels: the process group must outlive it) and its buffers."""
        if self._graph is not None:
            torch.cuda.synchronize(self.engine.device)
            self._graph = None
            self._graph_plan = None

    def phase_times(self):
        if not self.marks:
            return {}
        torch.cuda.synchronize()
        return {lab: a.elapsed_time(b) for (_, a), (lab, b) in zip(self.marks[:-1], self.marks[1:])}
