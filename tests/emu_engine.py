"""Drive the CPU-emulated build of the SIMT kernels (tests/emu) through the same ctypes binding the
product uses.  TEST INFRASTRUCTURE ONLY -- buffers are numpy arrays, "device pointers" are host pointers."""
from __future__ import annotations

import os
import subprocess
import sys
from typing import Dict, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlff_b200.capi import So3kLib  # noqa: E402

EMU_SO = os.path.join(ROOT, 'tests', 'emu', '_build', 'libso3k_emu.so')


def build_emu(force: bool = False) -> str:
    srcs = [os.path.join(ROOT, 'mlff_b200', 'csrc', f) for f in os.listdir(os.path.join(ROOT, 'mlff_b200', 'csrc'))]
    srcs.append(os.path.join(ROOT, 'tests', 'emu', 'cuda_emu.h'))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.check_call(['sh', os.path.join(ROOT, 'tests', 'emu', 'build_emu.sh')])
    return EMU_SO


def ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


class EmuEngine:
    def __init__(self, F, n_rbf, n_layers, n_heads, degrees, r_cut, flags=0, zbl_shift=0.0):
        self.lib = So3kLib(build_emu())
        self.h = self.lib.create(F, n_rbf, n_layers, n_heads, list(degrees), r_cut, 100, flags, zbl_shift)
        self.F, self.M = F, sum(2 * l + 1 for l in degrees)
        self.K = n_rbf

    def load(self, params: Dict[str, np.ndarray]):
        n = self.lib.param_count(self.h)
        blob = np.zeros(n, dtype=np.float32)
        used = 0
        for name, arr in params.items():
            off, cnt = self.lib.param_lookup(self.h, name)
            assert cnt == arr.size, (name, cnt, arr.shape)
            blob[off:off + cnt] = np.asarray(arr, dtype=np.float32).reshape(-1)
            used += cnt
        assert used == n, (used, n)
        self.blob = blob
        self.lib.load_params(self.h, ptr(blob), n)

    def energy_forces(self, R, z, idx_i, idx_j, cell=None, cell_offset=None, want_stress=False):
        R = np.ascontiguousarray(R, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.int32)
        ii = np.ascontiguousarray(idx_i, dtype=np.int32)
        jj = np.ascontiguousarray(idx_j, dtype=np.int32)
        B, n = z.shape
        P = ii.shape[1]
        cell = None if cell is None else np.ascontiguousarray(cell, dtype=np.float32)
        off = None if cell_offset is None else np.ascontiguousarray(cell_offset, dtype=np.int32)
        E = np.zeros(B, np.float32)
        F = np.zeros((B, n, 3), np.float32)
        ea = np.zeros((B, n), np.float32)
        nb = self.lib.workspace_bytes(self.h, B * n, B * P)
        ws = np.zeros(nb // 4 + 64, np.float32)
        st = np.zeros(1, np.int32)
        self.lib.energy_forces(self.h, B, n, P, ptr(R), ptr(z), ptr(ii), ptr(jj), ptr(cell), ptr(off), ptr(E), ptr(F),
                               ptr(ea), ptr(ws), ws.nbytes, ptr(st))
        if want_stress:
            stress = np.zeros((B, 3, 3), np.float32)
            self.lib.stress(self.h, B, n, P, ptr(ws), ws.nbytes, ptr(stress))
            return E, F, ea, int(st[0]), stress
        return E, F, ea, int(st[0])

    def potential(self, edges, z, centers, others, mask=None, energy_scale=1.0):
        edges = np.ascontiguousarray(edges, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.int32)
        ci = np.ascontiguousarray(centers, dtype=np.int32)
        oj = np.ascontiguousarray(others, dtype=np.int32)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        N, P = len(z), len(ci)
        ea = np.zeros(N, np.float32)
        dE = np.zeros((P, 3), np.float32)
        Fo = np.zeros((N, 3), np.float32)
        nb = self.lib.workspace_bytes(self.h, N, P)
        ws = np.zeros(nb // 4 + 64, np.float32)
        st = np.zeros(1, np.int32)
        self.lib.potential_displacements(self.h, N, P, ptr(edges), ptr(z), ptr(ci), ptr(oj), ptr(mk), energy_scale,
                                         ptr(ea), ptr(dE), ptr(Fo), ptr(ws), ws.nbytes, ptr(st))
        return ea, dE, Fo, int(st[0])

    def featurise(self, n, idx_i, idx_j, R=None, cell=None, cell_offset=None, disp=None):
        ii = np.ascontiguousarray(idx_i, dtype=np.int32)
        jj = np.ascontiguousarray(idx_j, dtype=np.int32)
        P = len(ii)
        R = None if R is None else np.ascontiguousarray(R, dtype=np.float32)
        cell = None if cell is None else np.ascontiguousarray(cell, dtype=np.float32)
        off = None if cell_offset is None else np.ascontiguousarray(cell_offset, dtype=np.int32)
        disp = None if disp is None else np.ascontiguousarray(disp, dtype=np.float32)
        out = {'rbf': np.zeros((P, self.K), np.float32), 'phi': np.zeros(P, np.float32), 'd': np.zeros(P, np.float32),
               'unit': np.zeros((P, 3), np.float32), 'sph': np.zeros((P, self.M), np.float32)}
        self.lib.featurise(self.h, n, P, ptr(R), ptr(ii), ptr(jj), ptr(cell), ptr(off), ptr(disp), ptr(out['rbf']),
                           ptr(out['phi']), ptr(out['d']), ptr(out['unit']), ptr(out['sph']))
        return out
