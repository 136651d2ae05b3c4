"""Batched neighbour lists for training sets.

`get_indices` mirrors `mlff/indexing/indices.py:15-84` (non-periodic branch): for every frame of a padded batch all
pairs (i, j) with 0 < d_ij <= r_cut and z_i * z_j != 0, in row-major order, padded with -1 to the longest list of the
batch.  The reference loops over the frames in Python with a dense n x n distance matrix per frame; here ONE launch of
the device cell list handles the whole batch: the frames are laid out along x, far enough apart that no pair crosses a
frame boundary, and the flat list is split back into frames."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .capi import NL_DENSE


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
        nv = int(count.item())
        if nv <= cap:
            break
        cap = int(1.25 * nv) + 16
        engine._status.zero_()
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
