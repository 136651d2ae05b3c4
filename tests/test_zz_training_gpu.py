"""Training seam on the B200 through the C ABI (so3k_energy_forces_vjp, so3k_loss_*, so3k_adam_step): the parameter
gradient of the reference's loss against torch double backward of the float64 oracle (itself pinned to the reference's
loss.py executed from source, tests/test_training.py), default model F = 132 / 3 layers / degrees 1-3."""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden

pytestmark = pytest.mark.gpu


def _problem(B=4, F=132, degrees=(1, 2, 3), heads=4, layers=3):
    from mlff_b200.config import So3kratesConfig
    eth = golden('ethanol_32.npz')
    R = np.concatenate([eth['R'][:B].astype(np.float64), np.zeros((B, 1, 3))], 1)
    z = np.concatenate([np.tile(eth['z'][None].astype(np.int64), (B, 1)), np.zeros((B, 1), np.int64)], 1)
    nl = o.get_indices(R, z, 5.0)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 5), np.int64)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 5), np.int64)], 1)
    rng = np.random.default_rng(23)
    E_true = rng.normal(size=(B, 1)) - 20.0
    F_true = rng.normal(size=(B, R.shape[1], 3))
    F_true[1, 2] = np.nan
    ocfg = o.OracleConfig(F=F, degrees=degrees, n_heads=heads, n_layers=layers)
    p = o.init_params(ocfg, 1, embed_scale=4.0, scale_shift=True)
    cfg = So3kratesConfig(F=F, n_layer=layers, degrees=degrees, n_heads=heads)
    return ocfg, cfg, p, R, z, ii, jj, E_true, F_true


def _engine(cfg, p, flags):
    from mlff_b200.engine import So3kEngine
    return So3kEngine(cfg, 'cuda:0', flags).load_params({k: np.asarray(v, np.float32) for k, v in p.items()})


def _check(eng, grad, ref, p, rtol):
    worst = 0.0
    for k in p:
        if k.startswith('energy/per_atom_'):
            continue
        off, cnt = eng.lib.param_lookup(eng.h, k)
        r = np.asarray(ref[k]).reshape(-1)
        err = np.abs(grad[off:off + cnt] - r).max()
        worst = max(worst, err / max(np.abs(r).max(), 1e-3))
        assert err <= rtol * max(np.abs(r).max(), 1e-3), (k, err, np.abs(r).max())
    return worst


@pytest.mark.parametrize('flags', [0, 3])
def test_loss_gradient_matches_oracle_on_device(flags):
    """flags = 0: fp32 first pass; 3: tcgen05 first pass (its split-bf16 E, F enter the cotangents only)."""
    from mlff_b200.training import Trainer
    ocfg, cfg, p, R, z, ii, jj, E_true, F_true = _problem()
    eng = _engine(cfg, p, flags)
    tr = Trainer(eng, {'energy': 0.01, 'force': 0.99})
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    loss, grad, E, F = tr.loss_and_grad(t(R), t(z), t(ii), t(jj), t(E_true), t(F_true))
    torch.cuda.synchronize()
    assert eng.check_status() == 0
    lo, og, Eo, Fo = o.loss_and_param_grads(ocfg, p, R, z, ii, jj, E_true, F_true, 0.01, 0.99)
    assert abs(float(loss[0]) - lo) <= 2e-5 * abs(lo)
    # tolerance: fp32 accumulation over ~300 edge rows; the tensor first pass adds its 1e-5 relative force error to Fbar
    worst = _check(eng, grad.cpu().numpy(), og, p, 1e-4 if flags == 0 else 3e-4)
    print(f'flags {flags}: loss {float(loss[0]):.6f} vs {lo:.6f}, worst relative gradient error {worst:.2e}')


def test_first_order_gradient_and_energy_on_device():
    ocfg, cfg, p, R, z, ii, jj, _, _ = _problem(B=3, F=64, degrees=(1, 2, 3, 4), heads=2, layers=2)
    eng = _engine(cfg, p, 0)
    lib, h, dev = eng.lib, eng.h, eng.device
    B, n = z.shape
    P = ii.shape[1]
    d = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dev, dt)
    Rd, zd, id_, jd = d(R, torch.float32), d(z, torch.int32), d(ii, torch.int32), d(jj, torch.int32)
    cb = d(np.array([1.0, 0.5, -2.0]), torch.float32)
    grad = torch.zeros(lib.param_count(h), device=dev)
    E = torch.zeros(B, device=dev)
    ws = torch.empty(lib.train_workspace_bytes(h, B * n, B * P) + 1024, dtype=torch.uint8, device=dev)
    lib.energy_forces_vjp(h, B, n, P, Rd.data_ptr(), zd.data_ptr(), id_.data_ptr(), jd.data_ptr(), None, None, cb.data_ptr(),
                          None, grad.data_ptr(), E.data_ptr(), ws.data_ptr(), ws.numel(), eng._status.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert eng.check_status() == 0
    names = [k for k in p if not k.startswith('energy/per_atom_')]
    pt = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=(k in names)) for k, v in p.items()}
    Es = [o.forward(ocfg, pt, z[b], ii[b], jj[b], torch.float64, R=torch.tensor(R[b])).sum() for b in range(B)]
    gs = torch.autograd.grad(Es[0] + 0.5 * Es[1] - 2.0 * Es[2], [pt[k] for k in names], allow_unused=True)
    ref = {k: (np.zeros(np.shape(p[k])) if g is None else g.numpy()) for k, g in zip(names, gs)}
    Eo = np.array([float(e.detach()) for e in Es])
    assert np.abs(E.cpu().numpy() - Eo).max() <= 1e-5 * np.abs(Eo).max()
    _check(eng, grad.cpu().numpy(), ref, p, 1e-4)


def test_training_steps_reduce_the_loss_on_device():
    from mlff_b200.training import Trainer, Optimizer
    ocfg, cfg, p, R, z, ii, jj, E_true, F_true = _problem(B=8)
    eng = _engine(cfg, p, 3)
    tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer(clip_by_global_norm=10.0).get(learning_rate=2e-4))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    args = (t(R), t(z), t(ii), t(jj), t(E_true), t(F_true))
    losses = [float(tr.train_step(*args)['loss']) for _ in range(6)]
    assert eng.check_status() == 0
    assert losses[-1] < losses[0] and tr.state.step == 6
    # the updated parameters are what the engine now evaluates: the loss of a fresh evaluation continues the sequence
    assert float(tr.loss_and_grad(*args)[0][0]) < losses[0]


def test_graph_replayed_training_step_equals_the_eager_one():
    """train_step(cuda_graph=True): the captured step (device-resident Adam count) reproduces the eager sequence of updates."""
    from mlff_b200.training import Trainer, Optimizer
    ocfg, cfg, p, R, z, ii, jj, E_true, F_true = _problem(B=4)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    args = (t(R), t(z), t(ii), t(jj), t(E_true), t(F_true))
    out = {}
    for mode in (False, True):
        eng = _engine(cfg, p, 0)
        tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer().get(learning_rate=2e-4))
        losses = [float(tr.train_step(*args, cuda_graph=mode)['loss']) for _ in range(4)]
        torch.cuda.synchronize()
        assert eng.check_status() == 0 and tr.state.step == 4 and int(tr.state.step_dev.item()) == 4
        out[mode] = (losses, eng._blob.clone())
        tr.close()
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-5)
    # atomics in the column sums make the gradient differ in the last bits between runs; Adam normalises the step size
    assert (out[True][1] - out[False][1]).abs().max() < 1e-5
    assert out[False][0][-1] < out[False][0][0]


def test_loss_gradient_with_all_optional_branches_on_device():
    """LayerNorm (pre / post) + both ResidualMLPs + ZBL at the default width: every parameter, the ten ZBL scalars included."""
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.training import Trainer
    allk = dict(layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True)
    _, _, _, R, z, ii, jj, E_true, F_true = _problem()
    ocfg = o.OracleConfig(F=132, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.3, **allk)
    p = o.init_params(ocfg, 2, embed_scale=4.0, scale_shift=True, readout_gain=0.2)
    cfg = So3kratesConfig(F=132, n_layer=2, zbl_repulsion=True, zbl_repulsion_shift=0.3, **allk)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    lo, og, _, _ = o.loss_and_param_grads(ocfg, p, R, z, ii, jj, E_true, F_true, 0.01, 0.99)
    for flags in (0, 3):
        eng = _engine(cfg, p, flags)
        tr = Trainer(eng, {'energy': 0.01, 'force': 0.99})
        loss, grad, _, _ = tr.loss_and_grad(t(R), t(z), t(ii), t(jj), t(E_true), t(F_true))
        torch.cuda.synchronize()
        assert eng.check_status() == 0
        assert abs(float(loss[0]) - lo) <= 5e-5 * abs(lo)
        worst = _check(eng, grad.cpu().numpy(), og, p, 1e-4 if flags == 0 else 5e-4)
        print(f'all branches + ZBL, flags {flags}: loss {float(loss[0]):.6f} vs {lo:.6f}, worst relative gradient error {worst:.2e}')


def test_loss_gradient_with_stress_target_on_device(solid):
    """Energy + force + stress loss on a periodic cell (several images per pair) through the Trainer, tensor-core first pass."""
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.training import Trainer
    pos, cell, zs = solid['R'][:40].astype(np.float64), solid['unit_cell'].astype(np.float64), solid['z'][:40].astype(np.int64)
    si, sj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 5.0)
    ocfg = o.OracleConfig(F=132, n_layers=2)
    p = o.init_params(ocfg, 6, embed_scale=4.0)
    cfg = So3kratesConfig(F=132, n_layer=2)
    rng = np.random.default_rng(3)
    E_true, F_true, S_true = rng.normal(size=(1, 1)), rng.normal(size=(1, 40, 3)), rng.normal(size=(1, 3, 3))
    w = {'energy': 0.01, 'force': 0.6, 'stress': 0.39}
    lo, og, _, _, _ = o.loss_and_param_grads(ocfg, p, pos[None], zs[None], si[None], sj[None], E_true, F_true, w['energy'],
                                             w['force'], cell=cell[None], cell_offset=S[None], S_true=S_true,
                                             w_stress=w['stress'])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    for flags in (0, 3):
        eng = _engine(cfg, p, flags)
        tr = Trainer(eng, w)
        loss, grad, _, _ = tr.loss_and_grad(t(pos[None]), t(zs[None]), t(si[None]), t(sj[None]), t(E_true), t(F_true),
                                            cell=t(cell[None]), cell_offset=t(S[None]), S_true=t(S_true))
        torch.cuda.synchronize()
        assert eng.check_status() == 0
        assert abs(float(loss[0]) - lo) <= 5e-5 * abs(lo)
        worst = _check(eng, grad.cpu().numpy(), og, p, 1e-4 if flags == 0 else 5e-4)
        print(f'stress target, flags {flags}: loss {float(loss[0]):.6f} vs {lo:.6f}, worst relative gradient error {worst:.2e}')

