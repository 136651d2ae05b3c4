"""The loss kernels of the training step (so3k_loss_sums / so3k_loss_cotangents, so3k_train.cu) checked the way the
reference checks its loss (tests/test_loss.py of the reference: test_mse_loss, test_masked_and_scaled_mse_loss,
test_gradients_masked_and_scaled_mse_loss -- plain mean, padding leaves it unchanged, NaN labels are excluded, and the same
for the gradient with respect to a scale in front of the prediction), through the CPU-emulated kernels."""
import numpy as np

from emu_engine import build_emu, ptr
from mlff_b200.capi import So3kLib


def _lib():
    return So3kLib(build_emu())


def _loss(lib, E, F, E_true, F_true, z, w_energy=0.0, w_force=1.0):
    """(loss, mse_E, mse_F), E_bar, F_bar, sums -- what one training step asks of the two kernels."""
    B, n = z.shape
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    E, F, E_true, F_true, zi = f32(E), f32(F), f32(E_true), f32(F_true), np.ascontiguousarray(z, np.int32)
    sums, ls = np.zeros(4, np.float64), np.zeros(3, np.float32)
    Eb, Fb = np.zeros(B, np.float32), np.zeros((B, n, 3), np.float32)
    lib.loss_sums(B, n, ptr(E), ptr(F), ptr(E_true), ptr(F_true), ptr(zi), ptr(sums))
    lib.loss_cotangents(B, n, ptr(E), ptr(F), ptr(E_true), ptr(F_true), ptr(zi), ptr(sums), w_energy, w_force, ptr(Eb),
                        ptr(Fb), ptr(ls))
    return ls, Eb, Fb, sums


def _data(seed=0, B=10, n=12):
    rng = np.random.default_rng(seed)
    return rng.uniform(size=(B, n, 3)), rng.uniform(size=(B, n, 3)), rng.uniform(size=B), rng.uniform(size=B)


def test_plain_mean_squared_error():
    lib = _lib()
    a, b, ea, eb = _data()
    z = np.ones(a.shape[:2], np.int32)
    ls, _, _, sums = _loss(lib, ea, a, eb, b, z, w_energy=0.25, w_force=0.75)
    mse_f, mse_e = np.mean((b - a) ** 2), np.mean((eb - ea) ** 2)
    assert sums[1] == a.shape[0] and sums[3] == a.size
    assert abs(ls[2] - mse_f) <= 1e-6 * mse_f and abs(ls[1] - mse_e) <= 1e-6 * mse_e
    assert abs(ls[0] - (0.25 * mse_e + 0.75 * mse_f)) <= 1e-6


def test_padding_and_nan_labels_leave_the_loss_unchanged():
    lib = _lib()
    a, b, ea, eb = _data()
    B, n = a.shape[:2]
    z = np.ones((B, n), np.int32)
    mse = np.mean((b - a) ** 2)
    # two padding structures (z = 0, zeros everywhere): a plain mean over the padded arrays would be smaller
    a_ = np.concatenate([a, np.zeros((2, n, 3))])
    b_ = np.concatenate([b, np.zeros((2, n, 3))])
    z_ = np.concatenate([z, np.zeros((2, n), np.int32)])
    assert np.mean((b_ - a_) ** 2) < mse
    ls, _, Fb, sums = _loss(lib, np.append(ea, [0, 0]), a_, np.append(eb, [np.nan, np.nan]), b_, z_)
    assert sums[3] == a.size and sums[1] == B
    assert abs(ls[2] - mse) <= 1e-6 * mse
    assert not Fb[B:].any()
    # two REAL structures whose labels are NaN: excluded from numerator and denominator (the plain mean is NaN)
    rng = np.random.default_rng(7)
    extra = rng.uniform(size=(1, n, 3))
    a_ = np.concatenate([extra, a, extra])
    b_ = np.concatenate([np.full((1, n, 3), np.nan), b, np.full((1, n, 3), np.nan)])
    z_ = np.ones((B + 2, n), np.int32)
    assert np.isnan(np.mean((b_ - a_) ** 2))
    ls, _, Fb, sums = _loss(lib, np.concatenate([[0.3], ea, [0.3]]), a_, np.concatenate([[np.nan], eb, [np.nan]]), b_, z_)
    assert sums[3] == a.size and sums[1] == B
    assert abs(ls[2] - mse) <= 1e-6 * mse
    assert np.isfinite(Fb).all() and not Fb[0].any() and not Fb[-1].any()
    # a single padding ATOM inside a real structure (node_mask of the force term)
    z1 = z.copy()
    z1[3, 5] = 0
    ls, _, Fb, sums = _loss(lib, ea, a, eb, b, z1)
    keep = np.ones((B, n), bool)
    keep[3, 5] = False
    assert sums[3] == 3 * keep.sum()
    assert abs(ls[2] - np.mean(((b - a) ** 2)[keep])) <= 1e-6 * mse
    assert not Fb[3, 5].any()


def test_gradient_with_respect_to_a_scale_of_the_prediction():
    """d/dp mse(p a, b) = <F_bar(p a), a>: the cotangents the kernels hand to the parameter-gradient sweep, with and without
    padding / NaN labels."""
    lib = _lib()
    a, b, ea, eb = _data()
    B, n = a.shape[:2]
    p = np.pi
    grad_gt = np.mean(2.0 * (p * a - b) * a)
    loss_gt = np.mean((p * a - b) ** 2)
    z = np.ones((B, n), np.int32)
    ls, _, Fb, _ = _loss(lib, ea, p * a, eb, b, z)
    assert abs(ls[0] - loss_gt) <= 1e-6 * loss_gt
    assert abs(float((Fb.astype(np.float64) * a).sum()) - grad_gt) <= 1e-5 * abs(grad_gt)
    rng = np.random.default_rng(3)
    extra = rng.uniform(size=(1, n, 3))
    a_ = np.concatenate([extra, a, np.zeros((1, n, 3))])
    b_ = np.concatenate([np.full((1, n, 3), np.nan), b, np.zeros((1, n, 3))])
    z_ = np.concatenate([np.ones((1, n), np.int32), z, np.zeros((1, n), np.int32)])
    ls, Eb, Fb, _ = _loss(lib, np.concatenate([[0.1], ea, [0.0]]), p * a_, np.concatenate([[0.2], eb, [0.0]]), b_, z_,
                          w_energy=0.0, w_force=1.0)
    assert abs(ls[0] - loss_gt) <= 1e-6 * loss_gt
    assert abs(float((Fb.astype(np.float64) * a_).sum()) - grad_gt) <= 1e-5 * abs(grad_gt)
    assert not Eb.any()                                        # zero energy weight: no energy cotangent
    # the energy term: d/dp mse(p E, E_true) = <E_bar, E>
    ls, Eb, _, _ = _loss(lib, p * ea, a, eb, b, z, w_energy=1.0, w_force=0.0)
    assert abs(float((Eb.astype(np.float64) * ea).sum()) - np.mean(2.0 * (p * ea - eb) * ea)) <= 1e-5
