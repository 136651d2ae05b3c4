"""Training seam (SURVEY 8(b) so3k_energy_forces_vjp, 8(e) data-parallel training) on the CPU: the oracle against the
reference's own loss code, and the CPU-emulated build of the training kernels (the same so3k_train.cu the B200 runs)
against the oracle -- first-order dE/dtheta, the second-order loss gradient, the loss sums / cotangents, the optimiser
update, the Trainer driver and its gloo world_size-2 data-parallel path."""
import os
import sys

import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from emu_engine import EmuEngine, build_emu, ptr
from mlff_b200.config import So3kratesConfig

SMALL = dict(F=32, n_rbf=8, degrees=(1, 2), n_heads=4, n_layers=2)


def _batch(B=2, seed=17, frames=(0, 7)):
    eth = golden('ethanol_32.npz')
    R = np.concatenate([eth['R'][list(frames)].astype(np.float64), np.zeros((B, 1, 3))], 1)
    z = np.concatenate([np.tile(eth['z'][None].astype(np.int64), (B, 1)), np.zeros((B, 1), np.int64)], 1)
    nl = o.get_indices(R, z, 5.0)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 3), np.int64)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 3), np.int64)], 1)
    rng = np.random.default_rng(seed)
    E_true = rng.normal(size=(B, 1)) - 25.0
    F_true = rng.normal(size=(B, R.shape[1], 3))
    F_true[0, 4] = np.nan
    return R, z, ii, jj, E_true, F_true


def _trainable(p):
    return [k for k in p if not k.startswith('energy/per_atom_')]


# ---- oracle vs the reference's loss.py executed from source ---------------------------------------------------------
def test_oracle_loss_and_gradient_match_reference_training_code():
    g = golden('reference_training.npz')
    cfg = o.OracleConfig(**SMALL)
    p = o.init_params(cfg, int(g['seed']), embed_scale=4.0, scale_shift=True)
    R, z, ii, jj = (g['in/' + k] for k in ('R', 'z', 'idx_i', 'idx_j'))
    wE, wF = float(g['w_energy']), float(g['w_force'])
    for tag in ('full', 'nan_energy'):
        loss, grads, E, F = o.loss_and_param_grads(cfg, p, R, z, ii, jj, g[f'{tag}/E_true'], g[f'{tag}/F_true'], wE, wF)
        assert abs(loss - float(g[f'{tag}/loss'])) <= 1e-7 * abs(float(g[f'{tag}/loss']))
    # projections of the gradient: central differences of the reference loss along seeded parameter directions
    loss, grads, _, _ = o.loss_and_param_grads(cfg, p, R, z, ii, jj, g['full/E_true'], g['full/F_true'], wE, wF)
    names = _trainable(p)
    proj = g['full/grad_projections']
    assert np.abs(proj).min() > 1e-4                   # the fixture is discriminating
    for k, ref in enumerate(proj):
        drng = np.random.default_rng(100 + k)
        v = {n: drng.normal(size=np.shape(p[n])) for n in names}
        nrm = np.sqrt(sum(float((a ** 2).sum()) for a in v.values()))
        got = sum(float((grads[n] * v[n]).sum()) for n in names) / nrm
        assert abs(got - ref) <= 2e-5 * max(abs(ref), 1e-2), (k, got, ref)


def test_oracle_loss_masks_and_denominators():
    """scaled_safe_masked_mse_loss (loss.py:13-29): NaN targets and padding atoms leave numerator AND denominator."""
    y = torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=torch.float64)
    t = torch.tensor([[1.5, float('nan')], [2.0, 0.0]], dtype=torch.float64)
    m = torch.tensor([[True, True], [True, False]])
    v = o.scaled_safe_masked_mse_loss(y, t, torch.ones(()), m)
    assert abs(float(v) - (0.25 + 1.0) / 2) < 1e-15
    assert float(o.scaled_safe_masked_mse_loss(y, t, torch.ones(()), torch.zeros_like(m))) == 0.0


# ---- emulated kernels vs oracle ----------------------------------------------------------------------------------------
def _emu(cfg, p):
    eng = EmuEngine(cfg.F, cfg.n_rbf, cfg.n_layers, cfg.n_heads, cfg.degrees, cfg.r_cut)
    eng.load(p)
    return eng


def _vjp(eng, R, z, ii, jj, E_bar, F_bar, cell=None, off=None, S_bar=None):
    lib, h = eng.lib, eng.h
    B, n = z.shape
    P = ii.shape[1]
    Rf, zi = np.ascontiguousarray(R, np.float32), np.ascontiguousarray(z, np.int32)
    i32, j32 = np.ascontiguousarray(ii, np.int32), np.ascontiguousarray(jj, np.int32)
    cell = None if cell is None else np.ascontiguousarray(cell, np.float32)
    off = None if off is None else np.ascontiguousarray(off, np.int32)
    ws = np.zeros(lib.train_workspace_bytes(h, B * n, B * P) // 4 + 64, np.float32)
    st = np.zeros(1, np.int32)
    grad, E = np.zeros(lib.param_count(h), np.float32), np.zeros(B, np.float32)
    Eb = None if E_bar is None else np.ascontiguousarray(E_bar, np.float32)
    Fb = None if F_bar is None else np.ascontiguousarray(F_bar, np.float32)
    Sb = None if S_bar is None else np.ascontiguousarray(S_bar, np.float32)
    lib.energy_forces_vjp(h, B, n, P, ptr(Rf), ptr(zi), ptr(i32), ptr(j32), ptr(cell), ptr(off), ptr(Eb), ptr(Fb),
                          ptr(grad), ptr(E), ptr(ws), ws.nbytes, ptr(st), S_bar=ptr(Sb))
    assert st[0] == 0
    return grad, E


def _compare_blob(eng, grad, ref, names, rtol):
    for k in names:
        off, cnt = eng.lib.param_lookup(eng.h, k)
        r = np.asarray(ref[k]).reshape(-1)
        err = np.abs(grad[off:off + cnt] - r).max()
        assert err <= rtol * max(np.abs(r).max(), 1e-3), (k, err, np.abs(r).max())
    for k in ('energy/per_atom_scale', 'energy/per_atom_shift'):
        off, cnt = eng.lib.param_lookup(eng.h, k)
        assert not grad[off:off + cnt].any()


def test_first_order_parameter_gradient_matches_oracle():
    """F_bar = NULL: grad = sum_b E_bar[b] dE_b/dtheta (torch autograd of the oracle's forward)."""
    cfg = o.OracleConfig(**SMALL)
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    R, z, ii, jj, _, _ = _batch()
    eng = _emu(cfg, p)
    cb = np.array([1.0, -0.5])
    grad, E = _vjp(eng, R, z, ii, jj, cb, None)
    names = _trainable(p)
    pt = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=(k in names)) for k, v in p.items()}
    Es = [o.forward(cfg, pt, z[b], ii[b], jj[b], torch.float64, R=torch.tensor(R[b])).sum() for b in range(2)]
    gs = torch.autograd.grad(cb[0] * Es[0] + cb[1] * Es[1], [pt[k] for k in names], allow_unused=True)
    ref = {k: (np.zeros(np.shape(p[k])) if g is None else g.numpy()) for k, g in zip(names, gs)}
    np.testing.assert_allclose(E, [float(e.detach()) for e in Es], rtol=1e-6)
    _compare_blob(eng, grad, ref, names, 2e-5)


@pytest.mark.parametrize('case', ['molecules', 'repeated_degrees'])
def test_loss_gradient_matches_oracle(case):
    """The second-order path: cotangents from so3k_loss_*, gradient from the dual-number sweep, against torch double
    backward of the oracle (loss.py:42-73 through F = -dE/dR)."""
    cfg = o.OracleConfig(**SMALL) if case == 'molecules' else o.OracleConfig(F=24, n_rbf=8, degrees=(1, 2, 2), n_heads=2,
                                                                             n_layers=2)
    p = o.init_params(cfg, 3, embed_scale=4.0, scale_shift=True)
    R, z, ii, jj, E_true, F_true = _batch()
    eng = _emu(cfg, p)
    lib = eng.lib
    loss, og, E, F = o.loss_and_param_grads(cfg, p, R, z, ii, jj, E_true, F_true, 0.01, 0.99)
    B, n = z.shape
    # E, F of the first pass: the emulated fp32 kernels
    Ee, Fe, _, st = eng.energy_forces(R, z, ii, jj)
    assert st == 0
    zi = np.ascontiguousarray(z, np.int32)
    Et, Ft = np.ascontiguousarray(E_true.reshape(-1), np.float32), np.ascontiguousarray(F_true, np.float32)
    sums, ls = np.zeros(4, np.float64), np.zeros(3, np.float32)
    Eb, Fb = np.zeros(B, np.float32), np.zeros((B, n, 3), np.float32)
    lib.loss_sums(B, n, ptr(Ee), ptr(Fe), ptr(Et), ptr(Ft), ptr(zi), ptr(sums))
    assert sums[1] == B and sums[3] == 3 * ((z != 0).sum() - 1)            # one real atom carries NaN force labels
    lib.loss_cotangents(B, n, ptr(Ee), ptr(Fe), ptr(Et), ptr(Ft), ptr(zi), ptr(sums), 0.01, 0.99, ptr(Eb), ptr(Fb), ptr(ls))
    assert abs(ls[0] - loss) <= 2e-5 * abs(loss)
    assert not Fb[:, -1].any() and not Fb[0, 4].any()
    grad, _ = _vjp(eng, R, z, ii, jj, Eb, Fb)
    _compare_blob(eng, grad, og, _trainable(p), 5e-5)


OPT_SETS = {'ln+post': dict(layer_normalization=True, final_layer=True),
            'res1+res2': dict(residual_mlp_1=True, residual_mlp_2=True),
            'all+zbl': dict(layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True,
                            zbl_repulsion=True, zbl_repulsion_shift=0.2)}


@pytest.mark.parametrize('name', list(OPT_SETS))
def test_loss_gradient_with_optional_branches(name):
    """LayerNorm (pre / post), both ResidualMLPs and the ZBL repulsion (its ten scalars included) in the training sweep:
    so3krates_layer.py:103-106, 149-174, mlp.py:30-72, observable.py:110-183 -- every parameter of every model variant the
    inference path supports has a gradient."""
    from common import engine_flags
    cfg = o.OracleConfig(**SMALL, **OPT_SETS[name])
    p = o.init_params(cfg, 3, embed_scale=4.0, scale_shift=True, readout_gain=0.2 if 'ln' in name or 'all' in name else 1.0)
    R, z, ii, jj, E_true, F_true = _batch()
    eng = EmuEngine(cfg.F, cfg.n_rbf, cfg.n_layers, cfg.n_heads, cfg.degrees, cfg.r_cut, flags=engine_flags(cfg),
                    zbl_shift=cfg.zbl_repulsion_shift)
    eng.load(p)
    lib = eng.lib
    loss, og, E, F = o.loss_and_param_grads(cfg, p, R, z, ii, jj, E_true, F_true, 0.01, 0.99)
    if cfg.zbl_repulsion:
        assert max(np.abs(og['zbl/' + k]).max() for k in ('a1', 'c2', 'p', 'd')) > 1e-4       # the fixture exercises them
    B, n = z.shape
    Ee, Fe, _, st = eng.energy_forces(R, z, ii, jj)
    assert st == 0
    zi = np.ascontiguousarray(z, np.int32)
    Et, Ft = np.ascontiguousarray(E_true.reshape(-1), np.float32), np.ascontiguousarray(F_true, np.float32)
    sums, ls = np.zeros(4, np.float64), np.zeros(3, np.float32)
    Eb, Fb = np.zeros(B, np.float32), np.zeros((B, n, 3), np.float32)
    lib.loss_sums(B, n, ptr(Ee), ptr(Fe), ptr(Et), ptr(Ft), ptr(zi), ptr(sums))
    lib.loss_cotangents(B, n, ptr(Ee), ptr(Fe), ptr(Et), ptr(Ft), ptr(zi), ptr(sums), 0.01, 0.99, ptr(Eb), ptr(Fb), ptr(ls))
    assert abs(ls[0] - loss) <= 2e-5 * abs(loss)
    grad, Esw = _vjp(eng, R, z, ii, jj, Eb, Fb)
    np.testing.assert_allclose(Esw, E.numpy().reshape(-1), rtol=2e-6)            # incl. the per-structure ZBL shift
    _compare_blob(eng, grad, og, _trainable(p), 5e-5)


def test_loss_gradient_periodic_cell_offsets():
    sol = golden('solid_0.npz')
    pos, cell = sol['R'][:24].astype(np.float64), sol['unit_cell'].astype(np.float64)
    zs = sol['z'][:24].astype(np.int64)
    si, sj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 4.0)
    cfg = o.OracleConfig(F=24, n_rbf=8, degrees=(1, 2), n_heads=2, n_layers=2, r_cut=4.0)
    p = o.init_params(cfg, 5, embed_scale=4.0)
    rng = np.random.default_rng(2)
    E_true, F_true = rng.normal(size=(1, 1)), rng.normal(size=(1, 24, 3))
    R, z, ii, jj = pos[None], zs[None], si[None], sj[None]
    names = _trainable(p)
    pt = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=(k in names)) for k, v in p.items()}
    Rt = torch.tensor(pos, requires_grad=True)
    e = o.forward(cfg, pt, zs, si, sj, torch.float64, R=Rt, cell=torch.tensor(cell), cell_offset=torch.tensor(S)).sum()
    (gR,) = torch.autograd.grad(e, Rt, create_graph=True)
    loss = 0.01 * (e - float(E_true[0, 0])) ** 2 + 0.99 * ((-gR - torch.tensor(F_true[0])) ** 2).mean()
    gs = torch.autograd.grad(loss, [pt[k] for k in names], allow_unused=True)
    ref = {k: (np.zeros(np.shape(p[k])) if g is None else g.numpy()) for k, g in zip(names, gs)}
    E_bar = np.array([0.02 * (float(e.detach()) - float(E_true[0, 0]))])
    F_bar = 0.99 * 2 * ((-gR).detach().numpy() - F_true[0]) / (24 * 3)
    eng = _emu(cfg, p)
    grad, _ = _vjp(eng, R, z, ii, jj, E_bar, F_bar[None], cell=cell[None], off=S[None])
    _compare_blob(eng, grad, ref, names, 5e-5)


def test_loss_gradient_with_stress_target():
    """Stress in the loss (the CLI's get_energy_force_stress_fn branch, mlff_train_so3krates.py:366-369): the stress cotangent
    enters the dual-number sweep as a strain direction (edge vectors move as r -> (1 + eps) r).  Periodic cell with several
    images per pair; one NaN stress label."""
    sol = golden('solid_0.npz')
    pos, cell = sol['R'][:24].astype(np.float64), sol['unit_cell'].astype(np.float64)
    zs = sol['z'][:24].astype(np.int64)
    si, sj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 4.0)
    cfg = o.OracleConfig(F=24, n_rbf=8, degrees=(1, 2), n_heads=2, n_layers=2, r_cut=4.0)
    p = o.init_params(cfg, 5, embed_scale=4.0)
    rng = np.random.default_rng(2)
    E_true, F_true, S_true = rng.normal(size=(1, 1)), rng.normal(size=(1, 24, 3)), rng.normal(size=(1, 3, 3))
    S_true[0, 1, 2] = np.nan
    wE, wF, wS = 0.01, 0.6, 0.39
    loss, og, E, F, St = o.loss_and_param_grads(cfg, p, pos[None], zs[None], si[None], sj[None], E_true, F_true, wE, wF,
                                                cell=cell[None], cell_offset=S[None], S_true=S_true, w_stress=wS)
    eng = _emu(cfg, p)
    # the engine's stress (so3k_stress) is the quantity in the loss
    Ee, Fe, _, st, Se = eng.energy_forces(pos[None], zs[None], si[None], sj[None], cell[None], S[None], want_stress=True)
    assert st == 0 and np.abs(Se - St.numpy()).max() <= 2e-5 * np.abs(St.numpy()).max()
    E_bar = 2 * wE * (E.numpy().reshape(-1) - E_true.reshape(-1))
    F_bar = 2 * wF * (F.numpy() - F_true) / F_true.size
    ok = ~np.isnan(S_true)
    S_bar = np.where(ok, 2 * wS * (St.numpy() - np.where(ok, S_true, 0.0)) / ok.sum(), 0.0)
    grad, _ = _vjp(eng, pos[None], zs[None], si[None], sj[None], E_bar, F_bar, cell=cell[None], off=S[None], S_bar=S_bar)
    _compare_blob(eng, grad, og, _trainable(p), 5e-5)
    # and through the Trainer (stress weight, NaN label masked out of numerator and denominator)
    from mlff_b200.training import Trainer
    te = EmuTorchEngine(cfg, p)
    tr = Trainer(te, {'energy': wE, 'force': wF, 'stress': wS})
    lt, gt, _, _ = tr.loss_and_grad(*_as_t(pos[None], zs[None], si[None], sj[None], E_true, F_true),
                                    cell=torch.as_tensor(cell[None]), cell_offset=torch.as_tensor(S[None]),
                                    S_true=torch.as_tensor(S_true))
    assert abs(float(lt[0]) - loss) <= 2e-5 * abs(loss) and lt.numel() == 4
    _compare_blob(eng, gt.numpy(), og, _trainable(p), 5e-5)


def test_adam_step_matches_optax_chain():
    """clip_by_global_norm -> scale_by_adam -> add_decayed_weights(mask) -> scale(-lr) (optimizer.py:64-97; optax's
    published update rules restated in numpy, float64)."""
    lib = EmuEngine(**{'F': 8, 'n_rbf': 4, 'n_layers': 1, 'n_heads': 2, 'degrees': (1,), 'r_cut': 5.0}).lib
    rng = np.random.default_rng(0)
    n = 1000
    p0 = rng.normal(size=n).astype(np.float32)
    mask = (rng.random(n) < 0.7).astype(np.uint8)
    frozen = np.zeros(n, np.uint8)
    frozen[:10] = 1
    lr, b1, b2, eps, er, wd, clip = 1e-2, 0.9, 0.999, 1e-8, 0.0, 0.05, 3.0
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    rp, rm, rv = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    ss = np.zeros(1, np.float64)
    for step in range(1, 4):
        g = (rng.normal(size=n) * (2.0 if step == 2 else 0.05)).astype(np.float32)
        lib.adam_step(n, ptr(p), ptr(g), ptr(m), ptr(v), ptr(frozen), ptr(mask), step, lr, b1, b2, eps, er, wd, clip, ptr(ss))
        gd = g.astype(np.float64)
        nrm = np.sqrt((gd ** 2).sum())
        assert abs(ss[0] - nrm ** 2) <= 1e-9 * nrm ** 2
        gd = gd * (clip / nrm if nrm > clip else 1.0)
        rm = b1 * rm + (1 - b1) * gd
        rv = b2 * rv + (1 - b2) * gd ** 2
        u = (rm / (1 - b1 ** step)) / (np.sqrt(rv / (1 - b2 ** step) + er) + eps) + wd * rp * mask
        keep = frozen == 0
        rp = np.where(keep, rp - lr * u, rp)
        rm, rv = np.where(keep, rm, 0.0), np.where(keep, rv, 0.0)
        np.testing.assert_allclose(p, rp, rtol=2e-5, atol=2e-6)
    assert (p[:10] == p0[:10]).all()
    # the device-resident update count (CUDA-graph replays): `step` is ignored, the call increments step_dev
    rng = np.random.default_rng(0)
    rng.normal(size=n); rng.random(n)
    p2, m2, v2 = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    sd = np.zeros(1, np.int64)
    for step in range(1, 4):
        g = (rng.normal(size=n) * (2.0 if step == 2 else 0.05)).astype(np.float32)
        lib.adam_step(n, ptr(p2), ptr(g), ptr(m2), ptr(v2), ptr(frozen), ptr(mask), 0, lr, b1, b2, eps, er, wd, clip, ptr(ss),
                      step_dev=ptr(sd))
    assert sd[0] == 3
    np.testing.assert_allclose(p2, p, rtol=1e-6, atol=1e-7)


# ---- the Trainer driver on an emulator-backed engine stand-in, single process and gloo world_size 2 -------------------
class EmuTorchEngine:
    """What mlff_b200.training needs of an engine (lib, h, cfg, device, _blob, _status, _stream, energy_forces), backed
    by the CPU-emulated kernels and torch CPU tensors.  TEST INFRASTRUCTURE."""

    def __init__(self, ocfg, p):
        from mlff_b200.capi import So3kLib
        from mlff_b200 import params as P
        self.lib = So3kLib(build_emu())
        self.cfg = So3kratesConfig(F=ocfg.F, n_rbf=ocfg.n_rbf, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees),
                                   n_heads=ocfg.n_heads, r_cut=ocfg.r_cut)
        self.h = self.lib.create(ocfg.F, ocfg.n_rbf, ocfg.n_layers, ocfg.n_heads, list(ocfg.degrees), ocfg.r_cut)
        self.device = torch.device('cpu')
        self._blob = torch.from_numpy(P.pack_blob(self.cfg, {k: np.asarray(v, np.float32) for k, v in p.items()}))
        self.lib.load_params(self.h, self._blob.data_ptr(), self._blob.numel())
        self._status = torch.zeros(1, dtype=torch.int32)

    def _stream(self):
        return 0

    def energy_forces(self, R, z, ii, jj, cell=None, off=None, want_stress=False):
        B, n = z.shape
        P = ii.shape[1]
        E, F = torch.zeros(B), torch.zeros(B, n, 3)
        ws = torch.zeros(self.lib.workspace_bytes(self.h, B * n, B * P) // 4 + 64)
        dp = lambda t: None if t is None else t.data_ptr()
        self.lib.energy_forces(self.h, B, n, P, dp(R), dp(z), dp(ii), dp(jj), dp(cell), dp(off), dp(E), dp(F), None,
                               dp(ws), ws.numel() * 4, dp(self._status))
        if want_stress:
            S = torch.zeros(B, 3, 3)
            self.lib.stress(self.h, B, n, P, dp(ws), ws.numel() * 4, dp(S))
            return E[:, None], F, None, S
        return E[:, None], F, None


def _trainer_problem():
    cfg = o.OracleConfig(**SMALL)
    p = o.init_params(cfg, 4, embed_scale=4.0, scale_shift=True)
    R, z, ii, jj, E_true, F_true = _batch(B=4, frames=(0, 7, 12, 20))
    return cfg, p, R, z, ii, jj, E_true, F_true


def _as_t(*arrays):
    return [torch.as_tensor(np.ascontiguousarray(a)) for a in arrays]


def test_trainer_step_reduces_the_loss_and_matches_oracle_gradient():
    from mlff_b200.training import Trainer, Optimizer
    cfg, p, R, z, ii, jj, E_true, F_true = _trainer_problem()
    eng = EmuTorchEngine(cfg, p)
    tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer(clip_by_global_norm=10.0).get(learning_rate=2e-4))
    tR, tz, ti, tj, tE, tF = _as_t(R, z, ii, jj, E_true, F_true)
    loss0, grad, _, _ = tr.loss_and_grad(tR, tz, ti, tj, tE, tF)
    l0 = float(loss0[0])
    lo, og, _, _ = o.loss_and_param_grads(cfg, p, R, z, ii, jj, E_true, F_true, 0.01, 0.99)
    assert abs(l0 - lo) <= 2e-5 * abs(lo)
    for k in _trainable(p):
        off, cnt = eng.lib.param_lookup(eng.h, k)
        r = og[k].reshape(-1)
        assert np.abs(grad[off:off + cnt].numpy() - r).max() <= 5e-5 * max(np.abs(r).max(), 1e-3), k
    ev = tr.eval_step(tR, tz, ti, tj, tE, tF)                              # valid_step_fn: same loss, no gradient
    assert abs(float(ev['loss']) - lo) <= 2e-5 * abs(lo)
    # per-target scales (variance scaling of the CLI): term = weight * scale * MSE, metric = plain MSE
    trs = Trainer(eng, {'energy': 0.01, 'force': 0.99}, scales={'energy': 3.0, 'force': 0.5})
    ls, gs_, _, _ = trs.loss_and_grad(tR, tz, ti, tj, tE, tF)
    assert abs(float(ls[0]) - (0.03 * float(loss0[1]) + 0.495 * float(loss0[2]))) <= 1e-5 * abs(float(ls[0]))
    assert abs(float(ls[1]) - float(loss0[1])) <= 1e-6 * abs(float(loss0[1]))
    with pytest.raises(NotImplementedError):
        Trainer(eng, {'energy': 1.0, 'partial_charge': 1.0})
    loss0, grad, _, _ = tr.loss_and_grad(tR, tz, ti, tj, tE, tF)           # (the buffers are shared between the trainers of an engine)
    blob0 = eng._blob.clone()
    gn = np.sqrt(sum(float((g ** 2).sum()) for g in og.values()))
    metrics = tr.train_step(tR, tz, ti, tj, tE, tF)
    assert abs(float(metrics['gradients_norm']) - gn) <= 5e-5 * gn          # train_metrics['gradients_norm'], run.py:49
    for _ in range(2):
        metrics = tr.train_step(tR, tz, ti, tj, tE, tF)
    assert tr.state.step == 3 and not torch.equal(blob0, eng._blob)
    assert float(metrics['loss']) < l0                                       # loss of the third step's parameters
    assert float(tr.loss_and_grad(tR, tz, ti, tj, tE, tF)[0][0]) < float(metrics['loss'])


def _dp_worker(rank, world, port, q):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), 'oracle'), os.path.dirname(here)]
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mlff_b200.training import Trainer, Optimizer
        cfg, p, R, z, ii, jj, E_true, F_true = _trainer_problem()
        eng = EmuTorchEngine(cfg, p)
        tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer().get(learning_rate=1e-3), group=True)
        sl = slice(rank * 2, rank * 2 + 2)                     # B / world structures per rank
        args = _as_t(R[sl], z[sl], ii[sl], jj[sl], E_true[sl], F_true[sl])
        loss, grad, _, _ = tr.loss_and_grad(*args)
        l0, g0 = float(loss[0]), grad.clone().numpy()
        tr.train_step(*args)
        if rank == 0:
            q.put((l0, g0, eng._blob.clone().numpy()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_over_gloo_world_size_2():
    """B = 4 structures split over two processes: all-reduced loss sums and gradient blob equal the single-process step on
    the whole batch (global denominators of scaled_safe_masked_mse_loss), and so do the updated parameters."""
    import torch.multiprocessing as mp
    from mlff_b200.training import Trainer, Optimizer
    build_emu()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    l_dp, g_dp, blob_dp = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    cfg, p, R, z, ii, jj, E_true, F_true = _trainer_problem()
    eng = EmuTorchEngine(cfg, p)
    tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer().get(learning_rate=1e-3))
    args = _as_t(R, z, ii, jj, E_true, F_true)
    loss, grad, _, _ = tr.loss_and_grad(*args)
    assert abs(l_dp - float(loss[0])) <= 1e-6 * abs(float(loss[0]))
    g1 = grad.clone().numpy()
    assert np.abs(g_dp - g1).max() <= 2e-5 * np.abs(g1).max()
    tr.train_step(*args)
    np.testing.assert_allclose(blob_dp, eng._blob.numpy(), rtol=0, atol=2e-5)


def test_train_state_schedules_and_polyak_average():
    """CustomTrainState.apply_gradients (mlff/training/train_state.py:59-101): the update of a step is scaled by
    lr_decay_fn(step) (the counter before the update) and by the plateau factor, `improved` feeds the plateau counter, and
    valid_params = step_size * new + (1 - step_size) * previous parameters -- against a numpy restatement of the chain."""
    from functools import partial
    from mlff_b200.training import TrainState, Optimizer, lr_decay_exponential, lr_decay_on_plateau
    cfg, p, *_ = _trainer_problem()
    eng = EmuTorchEngine(cfg, p)
    n = eng._blob.numel()
    lr, ss = 1e-2, 0.3
    tx = Optimizer().get(learning_rate=lr)
    st = TrainState(eng, tx, lr_decay_fn=partial(lr_decay_exponential, transition_steps=2, decay_factor=0.5),
                    reduce_lr_on_plateau_fn=partial(lr_decay_on_plateau, patience=2, decay_factor=0.1), polyak_step_size=ss)
    assert st.scheduled
    frozen = st.frozen.numpy().astype(bool)
    rng = np.random.default_rng(0)
    rp = eng._blob.numpy().astype(np.float64).copy()
    rm, rv = np.zeros(n), np.zeros(n)
    b1, b2, eps = tx.b1, tx.b2, tx.eps
    plateau_length, plateau_count = 0, 0
    improved = [True, False, False, False, True]                 # the plateau of two bad epochs completes before step 4
    for k, imp in enumerate(improved):
        g = (rng.normal(size=n) * 0.1).astype(np.float32)
        cd, plateau_length, inc = lr_decay_on_plateau(plateau_count, plateau_length, patience=2, decay_factor=0.1)
        plateau_count += inc
        factor = cd * 0.5 ** (k / 2)
        old = rp.copy()
        gd = g.astype(np.float64)
        rm = b1 * rm + (1 - b1) * gd
        rv = b2 * rv + (1 - b2) * gd ** 2
        u = (rm / (1 - b1 ** (k + 1))) / (np.sqrt(rv / (1 - b2 ** (k + 1))) + eps)
        rp = np.where(frozen, rp, rp - lr * factor * u)
        st.apply_gradients(torch.from_numpy(g))
        np.testing.assert_allclose(st.params.numpy(), rp, rtol=3e-5, atol=3e-6, err_msg=str(k))
        np.testing.assert_allclose(st.valid_params.numpy(), ss * rp + (1 - ss) * old, rtol=3e-5, atol=3e-6, err_msg=str(k))
        assert (st.plateau_length, st.plateau_count) == (plateau_length, plateau_count)
        st.improved(imp)
        plateau_length = 0 if imp else plateau_length + 1
    assert plateau_count == 1 and st.step == len(improved)
    # defaults: no schedule, valid_params IS params
    st0 = TrainState(eng, tx)
    assert not st0.scheduled and st0.valid_params is st0.params
    assert lr_decay_exponential(4, 2, 0.5) == 0.25 and lr_decay_on_plateau(2, 5, 2, 0.1) == (0.1 ** 2, 1, 2)
