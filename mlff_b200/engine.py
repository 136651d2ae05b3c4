"""Host-side owner of one libso3k handle: device parameter blob, workspace, and the calls into the C ABI.

PyTorch is plumbing here (device memory, the current CUDA stream); every computation happens in libso3k.so.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .capi import (STATUS_ASYMMETRIC, STATUS_BAD_Z, STATUS_OVERFLOW, STATUS_TC_TIMEOUT, NL_ASE, NL_DENSE, So3kError)
from .config import So3kratesConfig
from .params import pack_blob


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class So3kEngine:
    def __init__(self, cfg: So3kratesConfig, device: str = 'cuda:0', flags: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError('mlff_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.load()
        from .capi import FLAG_ZBL, FLAG_LAYER_NORM, FLAG_RESIDUAL_MLP_1, FLAG_RESIDUAL_MLP_2, FLAG_FINAL_LAYER
        if cfg.zbl_repulsion:
            flags |= FLAG_ZBL
        if cfg.layer_normalization:
            flags |= FLAG_LAYER_NORM | (FLAG_FINAL_LAYER if cfg.final_layer else 0)
        if cfg.residual_mlp_1:
            flags |= FLAG_RESIDUAL_MLP_1
        if cfg.residual_mlp_2:
            flags |= FLAG_RESIDUAL_MLP_2
        self.h = self.lib.create(cfg.F, cfg.n_rbf, cfg.n_layer, cfg.n_heads, list(cfg.degrees), cfg.r_cut,
                                 cfg.num_embeddings, flags, cfg.zbl_repulsion_shift)
        self._ws: Optional[torch.Tensor] = None
        self._nl_ws: Optional[torch.Tensor] = None
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._blob: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------------
    def load_params(self, params: Dict[str, np.ndarray]):
        blob = pack_blob(self.cfg, params)
        n = self.lib.param_count(self.h)
        if blob.size != n:
            raise So3kError(f'parameter blob has {blob.size} floats, the engine expects {n}')
        with torch.cuda.device(self.device):
            self._blob = torch.from_numpy(blob).to(self.device)
            self.lib.load_params(self.h, self._blob.data_ptr(), n, self._stream())
        return self

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, n_atoms: int, n_slots: int) -> torch.Tensor:
        nbytes = self.lib.workspace_bytes(self.h, n_atoms, n_slots)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def raise_for_status(s: int) -> int:
        """Raise So3kError for every ERROR bit of a status word already on the host (include/so3k.h, SO3K_STATUS_*).
        SO3K_STATUS_OVERFLOW is not an error: the caller re-allocates and repeats the call, exactly like
        PrimitiveNeighbors.overflow (mlff/indexing/indices.py:264-272); it is returned to the caller."""
        if s & STATUS_ASYMMETRIC:
            raise So3kError('edge list is not symmetric: an edge (i,j) has no reverse (j,i); outputs are undefined')
        if s & STATUS_BAD_Z:
            raise So3kError('atomic number outside [0, num_embeddings); outputs are undefined')
        if s & STATUS_TC_TIMEOUT:
            raise So3kError('a tensor-core kernel timed out waiting on an mbarrier; outputs are undefined')
        if s & ~(STATUS_OVERFLOW | STATUS_ASYMMETRIC | STATUS_BAD_Z | STATUS_TC_TIMEOUT):
            raise So3kError(f'unknown bits in the device status word: {s:#x}')
        return s

    def count_and_status(self, count: torch.Tensor) -> int:
        """ONE synchronising read of a neighbour-list edge count together with the device status word.  Error bits
        raised by any call since the last check (asymmetric list, bad Z, tensor-kernel time-out) raise here; the
        overflow bit is cleared (the caller compares the returned count with its capacity and re-allocates)."""
        v = torch.cat([count.reshape(1).to(torch.int64), self._status.to(torch.int64)]).cpu()
        self._status.zero_()
        self.raise_for_status(int(v[1]) & ~STATUS_OVERFLOW)
        return int(v[0])

    def check_status(self) -> int:
        """Synchronising read (and reset) of the device status word; raises on every error bit."""
        s = int(self._status.item())
        self._status.zero_()
        return self.raise_for_status(s)

    # ---- energy + forces (training / evaluation seam) ----------------------------------------------
    def energy_forces(self, R: torch.Tensor, z: torch.Tensor, idx_i: torch.Tensor, idx_j: torch.Tensor,
                      cell: Optional[torch.Tensor] = None, cell_offset: Optional[torch.Tensor] = None,
                      want_e_atom: bool = False, want_stress: bool = False):
        """R (B,n,3) f32, z (B,n) i32, idx (B,P) i32 [, cell (B,3,3) f32, cell_offset (B,P,3) i32] on the device.
        Returns E (B,1), F (B,n,3), e_atom (B,n) or None [, stress (B,3,3) = dE/d(strain) when want_stress:
        get_energy_force_stress_fn, observable_function.py:152-186]."""
        dev = self.device
        R = R.to(dev, torch.float32).contiguous()
        z = z.to(dev, torch.int32).contiguous()
        idx_i = idx_i.to(dev, torch.int32).contiguous()
        idx_j = idx_j.to(dev, torch.int32).contiguous()
        B, n = z.shape
        P = idx_i.shape[1]
        if cell is not None:
            cell = cell.to(dev, torch.float32).contiguous()
            cell_offset = cell_offset.to(dev, torch.int32).contiguous()
        E = torch.empty(B, dtype=torch.float32, device=dev)
        F = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
        ea = torch.empty(B, n, dtype=torch.float32, device=dev) if want_e_atom else None
        ws = self._workspace(B * n, B * P)
        with torch.cuda.device(dev):
            self.lib.energy_forces(self.h, B, n, P, _p(R), _p(z), _p(idx_i), _p(idx_j), _p(cell), _p(cell_offset),
                                   _p(E), _p(F), _p(ea), _p(ws), ws.numel(), _p(self._status), self._stream())
            if want_stress:
                st = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
                self.lib.stress(self.h, B, n, P, _p(ws), ws.numel(), _p(st), self._stream())
                return E[:, None], F, ea, st
        return E[:, None], F, ea

    def graphed_energy_forces(self, B: int, n: int, P: int, periodic: bool) -> 'GraphedEnergyForces':
        """One evaluation captured into a CUDA graph for fixed shapes (B, n, P): small systems are launch-bound (a step is
        ~75 kernel launches), a replay is one launch.  The library only enqueues on the given stream and never
        synchronises or allocates, so the capture needs no special path."""
        return GraphedEnergyForces(self, B, n, P, periodic)

    # ---- MD seam ------------------------------------------------------------------------------------
    def potential(self, edges: torch.Tensor, z: torch.Tensor, centers: torch.Tensor, others: torch.Tensor,
                  mask: Optional[torch.Tensor] = None, energy_scale: float = 1.0, want_dE_dedges: bool = True,
                  want_forces: bool = True):
        dev = self.device
        edges = edges.to(dev, torch.float32).contiguous()
        z = z.to(dev, torch.int32).contiguous()
        centers = centers.to(dev, torch.int32).contiguous()
        others = others.to(dev, torch.int32).contiguous()
        mk = None if mask is None else mask.to(dev, torch.uint8).contiguous()
        N, P = z.shape[0], centers.shape[0]
        ea = torch.empty(N, dtype=torch.float32, device=dev)
        dE = torch.empty(P, 3, dtype=torch.float32, device=dev) if want_dE_dedges else None
        Fo = torch.empty(N, 3, dtype=torch.float32, device=dev) if want_forces else None
        ws = self._workspace(N, P)
        with torch.cuda.device(dev):
            self.lib.potential_displacements(self.h, N, P, _p(edges), _p(z), _p(centers), _p(others), _p(mk),
                                             float(energy_scale), _p(ea), _p(dE), _p(Fo), _p(ws), ws.numel(),
                                             _p(self._status), self._stream())
        return ea, dE, Fo

    # ---- featurisation --------------------------------------------------------------------------------
    def featurise(self, idx_i: torch.Tensor, idx_j: torch.Tensor, n: int, R: Optional[torch.Tensor] = None,
                  cell: Optional[torch.Tensor] = None, cell_offset: Optional[torch.Tensor] = None,
                  displacements: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        dev = self.device
        idx_i = idx_i.to(dev, torch.int32).contiguous()
        idx_j = idx_j.to(dev, torch.int32).contiguous()
        P = idx_i.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        out = {'rbf_ij': torch.empty(P, self.cfg.n_rbf, **f32), 'phi_r_cut': torch.empty(P, **f32),
               'd_ij': torch.empty(P, **f32), 'unit_r_ij': torch.empty(P, 3, **f32),
               'sph_ij': torch.empty(P, self.cfg.M, **f32)}
        R = None if R is None else R.to(dev, torch.float32).contiguous()
        cell = None if cell is None else cell.to(dev, torch.float32).contiguous()
        cell_offset = None if cell_offset is None else cell_offset.to(dev, torch.int32).contiguous()
        disp = None if displacements is None else displacements.to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            self.lib.featurise(self.h, n, P, _p(R), _p(idx_i), _p(idx_j), _p(cell), _p(cell_offset), _p(disp),
                               _p(out['rbf_ij']), _p(out['phi_r_cut']), _p(out['d_ij']), _p(out['unit_r_ij']),
                               _p(out['sph_ij']), self._stream())
        return out

    # ---- neighbour list ---------------------------------------------------------------------------------
    def neighbor_list(self, positions: torch.Tensor, cutoff: float, capacity: int, cell: Optional[np.ndarray] = None,
                      pbc=(False, False, False), z: Optional[torch.Tensor] = None, mode: int = NL_ASE):
        """positions (N,3) float64 on the device; cell (3,3) float64 row-wise and pbc are HOST values.
        Returns idx_i, idx_j (capacity) int32, shifts (capacity,3) int32, count (device int64 scalar)."""
        dev = self.device
        pos = positions.to(dev, torch.float64).contiguous()
        N = pos.shape[0]
        zz = None if z is None else z.to(dev, torch.int32).contiguous()
        cell_h = None if cell is None else np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(3, 3))
        pbc_h = np.ascontiguousarray(np.asarray(pbc, dtype=np.uint8).reshape(3))
        i32 = dict(dtype=torch.int32, device=dev)
        ii = torch.empty(capacity, **i32)
        jj = torch.empty(capacity, **i32)
        S = torch.empty(capacity, 3, **i32)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        nb = self.lib.neighbor_list_workspace_bytes(N, capacity)
        if self._nl_ws is None or self._nl_ws.numel() < nb:
            self._nl_ws = torch.empty(int(nb * 1.05) + 1024, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            self.lib.neighbor_list(mode, N, _p(pos), _p(zz), None if cell_h is None else cell_h.ctypes.data,
                                   pbc_h.ctypes.data, cutoff, capacity, _p(ii), _p(jj), _p(S), _p(count),
                                   _p(self._nl_ws), self._nl_ws.numel(), _p(self._status), self._stream())
        return ii, jj, S, count


class GraphedEnergyForces:
    """Static buffers + a captured CUDA graph of So3kEngine.energy_forces for one shape.
    __call__(R, z, idx_i, idx_j[, cell, cell_offset]) copies the inputs into the static buffers and replays the graph;
    the returned E (B,1) and F (B,n,3) are the static output buffers (valid until the next call)."""

    def __init__(self, eng: So3kEngine, B: int, n: int, P: int, periodic: bool):
        dev = eng.device
        self.eng, self.shape, self.periodic = eng, (B, n, P), periodic
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        self.R = torch.zeros(B, n, 3, **f32)
        self.z = torch.zeros(B, n, **i32)
        self.ii = torch.full((B, P), -1, **i32)
        self.jj = torch.full((B, P), -1, **i32)
        self.cell = torch.zeros(B, 3, 3, **f32) if periodic else None
        self.off = torch.zeros(B, P, 3, **i32) if periodic else None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.E = self.F = None

    def _run(self):
        self.E, self.F, _ = self.eng.energy_forces(self.R, self.z, self.ii, self.jj, self.cell, self.off)

    def __call__(self, R, z, idx_i, idx_j, cell=None, cell_offset=None):
        self.R.copy_(R)
        self.z.copy_(z)
        self.ii.copy_(idx_i)
        self.jj.copy_(idx_j)
        if self.periodic:
            self.cell.copy_(cell)
            self.off.copy_(cell_offset)
        if self.graph is None:
            B, n, P = self.shape
            self.eng._workspace(B * n, B * P)                    # allocate outside the capture
            side = torch.cuda.Stream(device=self.eng.device)
            side.wait_stream(torch.cuda.current_stream(self.eng.device))
            with torch.cuda.stream(side):                         # warm-up: function attributes, lazy module loading
                self._run()
            torch.cuda.current_stream(self.eng.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
        self.graph.replay()
        return self.E, self.F
