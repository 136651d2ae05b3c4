"""Shared synthetic-input builders for the tests and bench (seeded numpy only; no reference data needed)."""
from __future__ import annotations

import numpy as np


def jittered_lattice(n_atoms: int, box: float, seed: int, jitter: float = 0.25, min_dist: float = 0.8):
    """Positions on a jittered simple-cubic lattice inside a cubic box (density as SURVEY 8(d): 0.1 atoms/A^3)."""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    a = box / m
    g = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing='ij'), -1).reshape(-1, 3)
    sel = rng.permutation(len(g))[:n_atoms]
    sel.sort()
    pos = (g[sel] + 0.5) * a + rng.uniform(-jitter, jitter, (n_atoms, 3)) * a
    return pos.astype(np.float64)


def water_like_numbers(n_atoms: int, seed: int):
    rng = np.random.default_rng(seed)
    return rng.choice(np.array([1, 6, 7, 8]), size=n_atoms, p=[0.62, 0.10, 0.03, 0.25]).astype(np.int64)


def as_sets(ii, jj, S=None):
    ii, jj = np.asarray(ii), np.asarray(jj)
    keep = ii >= 0
    if S is None:
        return sorted(zip(ii[keep].tolist(), jj[keep].tolist()))
    S = np.asarray(S)
    return sorted(zip(ii[keep].tolist(), jj[keep].tolist(), S[keep, 0].tolist(), S[keep, 1].tolist(), S[keep, 2].tolist()))


def engine_flags(cfg) -> int:
    """SO3K_FLAG_* bits of the optional model branches selected by an oracle / engine config."""
    from mlff_b200.capi import (FLAG_ZBL, FLAG_LAYER_NORM, FLAG_RESIDUAL_MLP_1, FLAG_RESIDUAL_MLP_2, FLAG_FINAL_LAYER)
    f = FLAG_ZBL if getattr(cfg, 'zbl_repulsion', False) else 0
    if getattr(cfg, 'layer_normalization', False):
        f |= FLAG_LAYER_NORM | (FLAG_FINAL_LAYER if getattr(cfg, 'final_layer', False) else 0)
    if getattr(cfg, 'residual_mlp_1', False):
        f |= FLAG_RESIDUAL_MLP_1
    if getattr(cfg, 'residual_mlp_2', False):
        f |= FLAG_RESIDUAL_MLP_2
    return f


def check_mic_list(oracle_mod, pos, cell, pbc, cutoff, ii, jj, S, cnt):
    """Device minimum-image list (SO3K_NL_MIC) against oracle.mic_neighbor_list (glp quadratic_neighbor_list +
    make_displacement semantics, mlff/mdx/atoms.py:488-529): the same ordered pairs, one image each, sorted by (i, j),
    and the image the shift selects reproduces the oracle's minimum-image displacement."""
    import numpy as np
    pos = np.asarray(pos, np.float64)
    cell_full = None if cell is None else np.asarray(cell, np.float64)
    oi, oj, odv = oracle_mod.mic_neighbor_list(pos, cell_full, cutoff, pbc=pbc)
    assert cnt == len(oi), (cnt, len(oi))
    ii, jj, S = ii[:cnt], jj[:cnt], S[:cnt]
    assert (ii == oi).all() and (jj == oj).all()
    dv = pos[jj] - pos[ii] + (S @ cell_full if cell_full is not None else 0.0)
    np.testing.assert_allclose(dv, odv, atol=1e-10)
    assert len(set(zip(ii.tolist(), jj.tolist()))) == cnt            # a single image per ordered pair
    return oi, oj, odv
