"""GPU parity: HBM-bound kernels vs the oracle.  Pool fwd/bwd (incl. tie routing) and the
confusion histogram are bit-exact; xent / Adam are fp32 with tolerance 1e-6 relative."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from tests.gpu_util import bf16_grid, dev_bf16, dev_f32, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


@pytest.mark.parametrize("shape,ties", [((2, 8, 12, 64), True), ((1, 160, 576, 64), False), ((3, 10, 36, 512), True),
                                        ((2, 6, 4, 8), True)])
def test_maxpool_fwd_bwd_bit_exact(ops, cuda_device, shape, ties):
    n, h, w, c = shape
    rng = np.random.default_rng(0)
    if ties:
        x = rng.integers(0, 3, shape).astype(np.float32)          # many ties, incl. post-ReLU zeros
    else:
        x = bf16_grid(np.maximum(rng.standard_normal(shape), 0))
    y_ref, idx_ref = T.max_pool_2x2_with_argmax(x)
    xd = dev_bf16(x, cuda_device)
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=cuda_device)
    idx = torch.empty((n, h // 2, w // 2, c), dtype=torch.uint8, device=cuda_device)
    ops.maxpool_fwd(xd, y, idx)
    torch.cuda.synchronize()
    assert np.array_equal(host(y), y_ref)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    dy = bf16_grid(rng.standard_normal(y_ref.shape))
    dx_ref = T.max_pool_2x2_grad(dy, idx_ref, (h, w))
    dx = torch.empty(shape, dtype=torch.bfloat16, device=cuda_device)
    ops.maxpool_bwd(dev_bf16(dy, cuda_device), idx, dx)
    torch.cuda.synchronize()
    assert np.array_equal(host(dx), dx_ref)
    # fused ReluGrad of the pre-pool activation
    ops.maxpool_bwd(dev_bf16(dy, cuda_device), idx, dx, act=xd)
    torch.cuda.synchronize()
    assert np.array_equal(host(dx), dx_ref * (x > 0))
    # the same mask read from the POOLED tensor (a quarter of the bytes): bit-identical for a ReLU output
    if x.min() >= 0:
        dx2 = torch.full(shape, 7.0, dtype=torch.bfloat16, device=cuda_device)
        ops.maxpool_bwd(dev_bf16(dy, cuda_device), idx, dx2, pooled=y)
        torch.cuda.synchronize()
        assert torch.equal(dx2, dx)
        # ... and BiasAddGrad of the conv in front of the pool from the same pass: column sums of dx
        db = torch.full((c,), 7.0, dtype=torch.float32, device=cuda_device)
        dx3 = torch.full(shape, 7.0, dtype=torch.bfloat16, device=cuda_device)
        ops.maxpool_bwd(dev_bf16(dy, cuda_device), idx, dx3, pooled=y, dbias=db)
        torch.cuda.synchronize()
        assert torch.equal(dx3, dx)
        ref_db = host(dx).astype(np.float64).sum(axis=(0, 1, 2))
        np.testing.assert_allclose(host(db), ref_db, rtol=1e-5, atol=1e-4 * np.abs(ref_db).max())


@pytest.mark.parametrize("npix", [1, 255, 4096, 2 * 160 * 576])
def test_softmax_xent_grad_argmax_confusion(ops, cuda_device, npix):
    rng = np.random.default_rng(1)
    lg = (rng.standard_normal((npix, 2)) * 3).astype(np.float32)
    lg[::7, 1] = lg[::7, 0]                                        # exact ties -> argmax 0
    lab = rng.integers(0, 2, npix).astype(np.uint8)
    onehot = np.eye(2, dtype=np.float32)[lab]
    lt = torch.tensor(lg, requires_grad=True)
    per = T.softmax_cross_entropy_with_logits(lt, torch.tensor(onehot))
    per.mean().backward()
    ld = dev_f32(lg, cuda_device)
    labd = torch.as_tensor(lab).to(cuda_device)
    dl = torch.empty_like(ld)
    pred = torch.empty(npix, dtype=torch.uint8, device=cuda_device)
    loss_sum = torch.zeros(2, dtype=torch.float32, device=cuda_device)
    cm = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    ws = ops.xent_workspace(npix, cuda_device)
    ops.softmax_xent(ld, labd, dl, pred, loss_sum, cm, ws, 1.0 / npix)
    torch.cuda.synchronize()
    assert abs(float(loss_sum[0]) / npix - float(per.mean())) <= 1e-5 * max(1.0, float(per.mean()))
    assert abs(float(loss_sum[1]) - float(loss_sum[0]) / npix) <= 1e-6 * max(1.0, float(loss_sum[1]))
    np.testing.assert_allclose(host(dl), lt.grad.numpy(), rtol=1e-5, atol=1e-7 / npix)
    pred_ref = T.argmax_last(torch.tensor(lg)).numpy().astype(np.uint8)
    assert np.array_equal(pred.cpu().numpy(), pred_ref)                       # bit-exact, ties -> 0
    cm_ref = T.confusion_matrix(lab, pred_ref)
    assert np.array_equal(cm.cpu().numpy().reshape(2, 2), cm_ref)             # bit-exact
    # deterministic: same bits twice
    loss2 = torch.zeros(2, dtype=torch.float32, device=cuda_device)
    ops.softmax_xent(ld, labd, None, None, loss2, None, ws, 1.0)
    torch.cuda.synchronize()
    assert float(loss2[0]) == float(loss_sum[0])


@pytest.mark.parametrize("npix", [16, 1000, 3 * 160 * 576 + 5])
def test_confusion_matrix_bit_exact(ops, cuda_device, npix):
    rng = np.random.default_rng(2)
    gt = rng.integers(0, 2, npix).astype(np.uint8)
    pr = rng.integers(0, 2, npix).astype(np.uint8)
    cm = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    ops.confusion_matrix(torch.as_tensor(gt).to(cuda_device), torch.as_tensor(pr).to(cuda_device), cm)
    ops.confusion_matrix(torch.as_tensor(gt).to(cuda_device), torch.as_tensor(pr).to(cuda_device), cm)   # accumulates
    torch.cuda.synchronize()
    assert np.array_equal(cm.cpu().numpy().reshape(2, 2), 2 * T.confusion_matrix(gt, pr))


def test_softmax_infer(ops, cuda_device):
    rng = np.random.default_rng(3)
    lg = (rng.standard_normal((1000, 2)) * 2).astype(np.float32)
    ld = dev_f32(lg, cuda_device)
    prob = torch.empty_like(ld)
    mask = torch.empty(1000, dtype=torch.uint8, device=cuda_device)
    ops.softmax_infer(ld, prob, mask)
    torch.cuda.synchronize()
    ref = torch.softmax(torch.tensor(lg), dim=-1).numpy()
    np.testing.assert_allclose(host(prob), ref, rtol=1e-5, atol=1e-7)
    assert np.array_equal(mask.cpu().numpy(), (lg[:, 1] > lg[:, 0]).astype(np.uint8))


def test_adam_tf_formula_three_steps(ops, cuda_device):
    rng = np.random.default_rng(4)
    n = 100003                                                          # odd tail
    p0 = rng.standard_normal(n).astype(np.float32)
    p0[::2] = 0.0                                                       # makes the update itself visible
    scales = 10.0 ** rng.uniform(-13, 0, n)                             # deep layers: |g| << eps
    p, m, v = torch.tensor(p0), torch.zeros(n), torch.zeros(n)
    pd = torch.zeros(n + 1, dtype=torch.float32, device=cuda_device)[:n]
    pd.copy_(torch.tensor(p0))
    md, vd = torch.zeros_like(pd), torch.zeros_like(pd)
    from semanticsegmentation_tensorflow_b200.plan import adam_lr_t
    gmax = np.zeros(n, np.float32)
    updmax = np.zeros(n, np.float32)
    for t in range(1, 4):
        g = (rng.standard_normal(n) * scales).astype(np.float32)
        gmax = np.maximum(gmax, np.abs(g))
        before = p.numpy().copy()
        T.adam_tf_step(p, m, v, torch.tensor(g), t)
        updmax = np.maximum(updmax, np.abs(p.numpy() - before))
        ops.adam_step(pd, md, vd, dev_f32(g, cuda_device), adam_lr_t(1e-4, t))
    torch.cuda.synchronize()
    # m can cancel (alternating gradient signs): tolerance relative to the gradient scale seen
    assert np.all(np.abs(host(md) - m.numpy()) <= 1e-6 * gmax + 1e-45)
    np.testing.assert_allclose(host(vd), v.numpy(), rtol=1e-5, atol=0)
    np.testing.assert_allclose(host(pd), p.numpy(), rtol=2e-7, atol=1e-9)
    # where p started at 0 the parameter IS the accumulated update (|g| << eps elements included):
    # error bounded relative to the largest single-step update of that element (m may cancel)
    assert np.all(np.abs(host(pd) - p.numpy())[::2] <= 1e-4 * updmax[::2] + 1e-30)


def test_momentum_step(ops, cuda_device):
    rng = np.random.default_rng(5)
    n = 4099
    p0 = rng.standard_normal(n).astype(np.float32)
    p, a = torch.tensor(p0), torch.zeros(n)
    pd, ad = dev_f32(p0, cuda_device), torch.zeros(n, device=cuda_device)
    for _ in range(3):
        g = rng.standard_normal(n).astype(np.float32)
        T.momentum_tf_step(p, a, torch.tensor(g), 0.01, 0.9)
        ops.momentum_step(pd, ad, dev_f32(g, cuda_device), 0.01, 0.9)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(pd), p.numpy(), rtol=1e-6, atol=1e-7)


def test_dropout_injected_mask_and_philox_rate(ops, cuda_device):
    rng = np.random.default_rng(6)
    n = 1 << 20
    x = bf16_grid(rng.standard_normal(n))
    mask = (rng.random(n) < 0.8).astype(np.uint8)
    xd = dev_bf16(x, cuda_device)
    y = torch.empty_like(xd)
    ops.dropout(xd, y, 0.8, 0, torch.as_tensor(mask).to(cuda_device))
    torch.cuda.synchronize()
    ref = bf16_grid(T.dropout(torch.tensor(x), 0.8, torch.tensor(mask.astype(np.float32))).numpy())
    assert np.array_equal(host(y), ref)
    ops.dropout(xd, y, 0.8, 1234)
    y2 = torch.empty_like(xd)
    ops.dropout(xd, y2, 0.8, 1234)
    y3 = torch.empty_like(xd)
    ops.dropout(xd, y3, 0.8, 1235)
    torch.cuda.synchronize()
    kept = (host(y) != 0) | (x == 0)
    assert abs(kept.mean() - 0.8) < 5e-3                                  # Bernoulli(keep) rate
    assert np.array_equal(host(y), host(y2))                             # same seed -> same mask (bwd regen)
    assert not np.array_equal(host(y), host(y3))
    nz = host(y) != 0
    assert np.array_equal(host(y)[nz], bf16_grid(x[nz] / np.float32(0.8)))
    # the 8-wide form (n % 8 == 0) and the scalar form (here: n - 4 elements) draw the same Philox stream
    ys = torch.empty(n - 4, dtype=torch.bfloat16, device=cuda_device)
    ops.dropout(xd[:n - 4], ys, 0.8, 1234)
    torch.cuda.synchronize()
    assert torch.equal(ys, y[:n - 4])
    ms = torch.as_tensor(mask).to(cuda_device)
    ops.dropout(xd[:n - 4], ys, 0.8, 0, ms[:n - 4])
    torch.cuda.synchronize()
    assert np.array_equal(host(ys), ref[:n - 4])


def test_bias_grad_and_cast(ops, cuda_device):
    rng = np.random.default_rng(7)
    dy = bf16_grid(rng.standard_normal((5000, 64)))
    db = torch.empty(64, dtype=torch.float32, device=cuda_device)
    ops.bias_grad(dev_bf16(dy, cuda_device), db)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(db), dy.sum(0), rtol=1e-4, atol=1e-3)
    dyf = rng.standard_normal((3000, 2)).astype(np.float32)
    db2 = torch.empty(2, dtype=torch.float32, device=cuda_device)
    ops.bias_grad(dev_f32(dyf, cuda_device), db2)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(db2), dyf.sum(0), rtol=1e-4, atol=1e-3)
    img = rng.integers(0, 256, (2, 8, 8, 3), dtype=np.uint8)
    out = torch.empty((2, 8, 8, 3), dtype=torch.bfloat16, device=cuda_device)
    ops.cast_to_bf16(torch.as_tensor(img).to(cuda_device), out)
    torch.cuda.synchronize()
    assert np.array_equal(host(out), img.astype(np.float32))


@pytest.mark.parametrize("c", [3, 4])
def test_overlay_mask_bit_exact(ops, cuda_device, c):
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (2, 37, 53, c), dtype=np.uint8)
    prob = rng.random((2, 37, 53)).astype(np.float32)
    mask = (prob > 0.5).astype(np.uint8)
    out = ops.overlay_mask(torch.as_tensor(img).to(cuda_device), torch.as_tensor(mask).to(cuda_device))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), T.paste_mask(img, prob))


@pytest.mark.parametrize("shape", [(3, 3, 64, 128), (1, 1, 192, 64), (7, 7, 64, 64), (3, 3, 32, 96), (3, 3, 3, 64)])
def test_pack_conv_weights_layouts_bit_exact(ops, cuda_device, shape):
    """fp32 HWIO master (FCN.py:125) -> wk[tap][Cin/64][Cout][64] and wd[rot180 tap][Cout/64][Cin][64]: the
    round-to-nearest-even bf16 of the master in the blocked kernel layout, zero-padded in the k dimension."""
    kh, kw, ci, co = shape
    rng = np.random.default_rng(40)
    w = rng.standard_normal(shape).astype(np.float32)
    wd_ = torch.as_tensor(w).to(cuda_device)
    wk, wd = ops.pack_conv_weights(wd_)
    torch.cuda.synchronize()
    ref = torch.as_tensor(w).to(torch.bfloat16).float().numpy().reshape(kh * kw, ci, co)

    def blocked(m):      # [T][rows][K] -> [T][ceil(K/64)][rows][64]
        t, rows, k = m.shape
        kc = -(-k // 64)
        p = np.zeros((t, rows, kc * 64), np.float32)
        p[:, :, :k] = m
        return p.reshape(t, rows, kc, 64).transpose(0, 2, 1, 3)

    assert np.array_equal(wk.float().cpu().numpy(), blocked(ref.transpose(0, 2, 1)))
    assert np.array_equal(wd.float().cpu().numpy(), blocked(ref[::-1]))


def test_onehot_to_ids_and_argmax(ops, cuda_device):
    rng = np.random.default_rng(9)
    ids = rng.integers(0, 5, (3, 7, 11)).astype(np.uint8)
    oh = np.eye(5, dtype=np.float32)[ids]
    out = torch.empty(ids.shape, dtype=torch.uint8, device=cuda_device)
    for t in (torch.as_tensor(oh), torch.as_tensor(oh.astype(np.uint8)), torch.as_tensor(oh.astype(bool))):
        ops.onehot_to_ids(t.to(cuda_device), out.zero_())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ids)
    # tf.argmax semantics of segk_softmax_infer: first index on ties
    lg = torch.as_tensor(rng.integers(0, 3, (1000, 5)).astype(np.float32)).to(cuda_device)     # many exact ties
    am = torch.empty(1000, dtype=torch.uint8, device=cuda_device)
    ops.softmax_infer(lg, None, None, am)
    torch.cuda.synchronize()
    assert np.array_equal(am.cpu().numpy(), lg.cpu().numpy().argmax(-1))


def test_adam_fused_with_repack_and_multi_range(ops, cuda_device):
    """segk_adam_pack_conv_weights == segk_adam_step followed by segk_pack_conv_weights, bit for bit; and
    segk_adam_step_ranges == segk_adam_step on each range (the rest of the arena untouched)."""
    g_ = torch.Generator(device=cuda_device).manual_seed(8)
    shape = (3, 3, 128, 192)
    mk = lambda s=1.0: torch.randn(shape, generator=g_, device=cuda_device) * s
    p, m, v, g = mk(0.05), mk(1e-3), mk(1e-3).abs() * 1e-3, mk(1e-2)
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    lr_t = 3.1e-4
    ops.adam_step(p.view(-1), m.view(-1), v.view(-1), g.view(-1), lr_t)
    wk, wd = ops.pack_conv_weights(p)
    wk2, wd2 = torch.zeros_like(wk), torch.zeros_like(wd)
    ops.adam_pack_conv_weights(p2, m2, v2, g, wk2, wd2, lr_t)
    torch.cuda.synchronize()
    assert torch.equal(p, p2) and torch.equal(m, m2) and torch.equal(v, v2)
    assert torch.equal(wk, wk2) and torch.equal(wd, wd2)
    # with the folded BN scale: == adam_step, scale_columns (gamma * mult), pack_conv_weights; master weights unscaled
    gamma = 1 + 0.2 * torch.randn(shape[3], generator=g_, device=cuda_device)
    p3, m3, v3 = p2.clone(), m2.clone(), v2.clone()
    ops.adam_step(p2.view(-1), m2.view(-1), v2.view(-1), g.view(-1), lr_t)
    wk, wd = ops.pack_conv_weights(ops.scale_columns(p2, gamma, 0.9995))
    wk3, wd3 = torch.zeros_like(wk), torch.zeros_like(wd)
    ops.adam_pack_conv_weights(p3, m3, v3, g, wk3, wd3, lr_t, col_scale=gamma, col_mult=0.9995)
    torch.cuda.synchronize()
    assert torch.equal(p2, p3) and torch.equal(m2, m3) and torch.equal(v2, v3)
    assert torch.equal(wk, wk3) and torch.equal(wd, wd3)
    n = 10000
    P_, M_, V_, G_ = (torch.randn(n, generator=g_, device=cuda_device) for _ in range(4))
    V_.abs_()
    ref = [t.clone() for t in (P_, M_, V_)]
    ranges = [(0, 7), (64, 100), (1024, 4097), (9000, 1000)]
    ops.adam_step_ranges(P_, M_, V_, G_, ranges, lr_t)
    for o, l in ranges:
        ops.adam_step(ref[0][o:o + l], ref[1][o:o + l], ref[2][o:o + l], G_[o:o + l], lr_t)
    torch.cuda.synchronize()
    for a, b in zip((P_, M_, V_), ref):
        assert torch.allclose(a, b, rtol=1e-6, atol=0)
