"""Batched neighbour lists for training sets.

`get_indices` mirrors `mlff/indexing/indices.py:15-84` (non-periodic branch): for every frame of a padded batch all
pairs (i, j) with 0 < d_ij <= r_cut and z_i * z_j != 0, in row-major order, padded with -1 to the longest list of the
batch.  The reference loops over the frames in Python with a dense n x n distance matrix per frame; here ONE launch of
the device cell list handles the whole batch: the frames are laid out along x, far enough apart that no pair crosses a
frame boundary, and the flat list is split back into frames.

`get_pbc_neighbors` mirrors `mlff/indexing/indices.py:111-191` (periodic training sets): per frame the ASE
`primitive_neighbor_list('ijS')` edge set of the real atoms (`pos[node_mask]`) for that frame's cell and pbc, every
frame padded to the longest list with -1 indices / zero shifts.  The reference's allocate / update / overflow
book-keeping only exists to bound its padded length and yields the same arrays; here each frame is one device
cell-list launch.  Edge order within a frame: sorted by (i, j, S) (ASE's traversal order differs; the SET is
identical -- the neighbour-list parity tests compare sorted sets).

`pad_*` mirror `mlff/padding/padding.py:47-108` and `mlff/indexing/indices.py:87-108` on numpy arrays or torch tensors."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .capi import NL_ASE, NL_DENSE



def _count(engine, count) -> int:
    """Edge count of a neighbour-list call, read together with the device status word (So3kEngine.count_and_status:
    error bits raise, the overflow bit is cleared -- the callers below compare the count with their capacity)."""
    f = getattr(engine, 'count_and_status', None)
    if f is not None:
        return f(count)
    n = int(count.item())
    if hasattr(engine, '_status'):
        engine._status.zero_()
    return n


def get_indices(engine, R, z, r_cut: float) -> Dict[str, torch.Tensor]:
    """R (B,n,3), z (B,n) (0 = padding atom), numpy or torch.  Returns {'idx_i', 'idx_j'}: (B,P_max) int32 on the engine's
    device, per-frame atom indices, -1 padded."""
    dev = engine.device
    R = torch.as_tensor(R).to(dev, torch.float64)
    z = torch.as_tensor(z).to(dev, torch.int32)
    B, n = int(z.shape[0]), int(z.shape[1])
    # frame b is shifted by b * pitch along x; pitch > extent of a frame + r_cut keeps the frames apart
    span = float((R[..., 0].amax() - R[..., 0].amin()).item()) if R.numel() else 0.0
    pitch = span + 2.0 * r_cut + 1.0
    off = torch.zeros(B, 1, 3, dtype=torch.float64, device=dev)
    off[:, 0, 0] = torch.arange(B, dtype=torch.float64, device=dev) * pitch
    pos = (R + off).reshape(B * n, 3)
    cap = max(64, B * n * 32)
    while True:
        ii, jj, _, count = engine.neighbor_list(pos, r_cut, cap, z=z.reshape(-1), mode=NL_DENSE)
        nv = _count(engine, count)
        if nv <= cap:
            break
        cap = int(1.25 * nv) + 16
    ii, jj = ii[:nv].long(), jj[:nv].long()                 # flat indices, sorted by (i, j): row-major inside a frame
    frame = ii // n
    per_frame = torch.bincount(frame, minlength=B)
    pmax = int(per_frame.max().item()) if nv else 0
    start = torch.cumsum(per_frame, 0) - per_frame
    slot = torch.arange(nv, device=dev) - start[frame]
    idx_i = torch.full((B, pmax), -1, dtype=torch.int32, device=dev)
    idx_j = torch.full((B, pmax), -1, dtype=torch.int32, device=dev)
    idx_i[frame, slot] = (ii - frame * n).to(torch.int32)
    idx_j[frame, slot] = (jj - frame * n).to(torch.int32)
    return {'idx_i': idx_i, 'idx_j': idx_j}


def get_pbc_neighbors(engine, pos, node_mask, pbc, cell, cutoff: float) -> Dict[str, torch.Tensor]:
    """pos (B,n,3), node_mask (B,n) bool (padding atoms False, at the end of a frame as the reference pads them),
    pbc (B,3), cell (B,3,3) row-wise (ASE default), numpy or torch.  Returns {'idx_i', 'idx_j'} (B,P_max) int32 and
    'shifts' (B,P_max,3) int32 on the engine's device."""
    dev = engine.device
    pos = torch.as_tensor(pos).to(dev, torch.float64)
    mask = torch.as_tensor(np.asarray(node_mask.cpu() if torch.is_tensor(node_mask) else node_mask)).to(torch.bool)
    cell_h = np.asarray(cell.cpu() if torch.is_tensor(cell) else cell, dtype=np.float64)
    pbc_h = np.asarray(pbc.cpu() if torch.is_tensor(pbc) else pbc).astype(bool)
    B = int(pos.shape[0])
    frames = []
    cap = 0
    for b in range(B):
        n_real = int(mask[b].sum())
        if not bool(mask[b, :n_real].all()):
            raise ValueError('node_mask: padding atoms must follow the real atoms of a frame')
        if n_real == 0:                                     # an all-padding frame has no edges
            e = torch.empty(0, dtype=torch.int32, device=dev)
            frames.append((e, e, torch.empty(0, 3, dtype=torch.int32, device=dev)))
            continue
        p = pos[b, :n_real]
        cap = max(cap, 64 * n_real + 16)
        while True:
            ii, jj, S, count = engine.neighbor_list(p, cutoff, cap, cell=cell_h[b], pbc=tuple(pbc_h[b]), mode=NL_ASE)
            nv = _count(engine, count)
            if nv <= cap:
                break
            cap = int(1.25 * nv) + 16                       # overflow: re-allocate, as the reference does (:151-155)
        frames.append((ii[:nv].clone(), jj[:nv].clone(), S[:nv].clone()))
    pmax = max((len(f[0]) for f in frames), default=0)
    idx_i = torch.full((B, pmax), -1, dtype=torch.int32, device=dev)
    idx_j = torch.full((B, pmax), -1, dtype=torch.int32, device=dev)
    shifts = torch.zeros((B, pmax, 3), dtype=torch.int32, device=dev)
    for b, (ii, jj, S) in enumerate(frames):
        idx_i[b, :len(ii)], idx_j[b, :len(ii)], shifts[b, :len(ii)] = ii, jj, S
    return {'idx_i': idx_i, 'idx_j': idx_j, 'shifts': shifts}


# ---- padding of a batch to common lengths (mlff/padding/padding.py) -------------------------------------------------
def _pad_axis(a, axis: int, target: int, pad_value):
    n = a.shape[axis]
    assert target - n >= 0
    if torch.is_tensor(a):
        shape = list(a.shape)
        shape[axis] = target - n
        return torch.cat([a, torch.full(shape, pad_value, dtype=a.dtype, device=a.device)], dim=axis)
    width = [(0, 0)] * a.ndim
    width[axis] = (0, target - n)
    return np.pad(a, width, mode='constant', constant_values=pad_value)


def pad_atomic_types(z, n_max: int, pad_value=0):
    """z (B,n) -> (B,n_max); padding atoms carry z = 0 (they are masked everywhere, stacknet.py:130)."""
    return _pad_axis(z, -1, n_max, pad_value)


def pad_coordinates(R, n_max: int, pad_value=0):
    """R (B,n,3) -> (B,n_max,3)"""
    return _pad_axis(R, -2, n_max, pad_value)


def pad_forces(F, n_max: int, pad_value=0):
    """F (B,n,3) -> (B,n_max,3)"""
    return _pad_axis(F, -2, n_max, pad_value)


def pad_per_atom_quantity(q, n_max: int, pad_value=0):
    """q (B,n,k) -> (B,n_max,k)"""
    return _pad_axis(q, -2, n_max, pad_value)


def pad_indices(idx_i, idx_j, n_pair_max: int, pad_value=-1):
    """idx_i, idx_j (B,P) -> (B,n_pair_max) each; padding slots carry -1 (pair_mask, stacknet.py:131)."""
    assert idx_i.shape[-1] == idx_j.shape[-1]
    return _pad_axis(idx_i, -1, n_pair_max, pad_value), _pad_axis(idx_j, -1, n_pair_max, pad_value)


def pad_index_list(idx, n_pair_max: int, pad_value=-1):
    """idx (P) -> (n_pair_max)                                                   mlff/indexing/indices.py:87-96"""
    return _pad_axis(idx, -1, n_pair_max, pad_value)


def pad_shift(shift, n_pair_max: int, pad_value=0):
    """shift (P,3) -> (n_pair_max,3)                                             mlff/indexing/indices.py:99-108"""
    return _pad_axis(shift, 0, n_pair_max, pad_value)
