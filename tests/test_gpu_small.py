"""GPU parity: CUDA-core kernels for the ragged-channel layers (conv1_1, conv8, conv_t1,
conv_t3) vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from tests.gpu_util import assert_close, bf16_grid, dev_bf16, dev_f32, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


@pytest.mark.parametrize("cin", [3, 4])
def test_conv1_1_u8_fwd_and_wgrad(ops, cuda_device, cin):
    n, h, w, co = 2, 16, 24, 64
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (n, h, w, cin), dtype=np.uint8)
    wt = (rng.standard_normal((3, 3, cin, co)) * 0.01).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    wtt = torch.tensor(wt, requires_grad=True)
    z = T.bias_add(T.conv2d_same(torch.tensor(img.astype(np.float32)), wtt), torch.tensor(b))
    y_ref = T.relu(z)
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    imgd = torch.as_tensor(img).to(cuda_device)
    ops.conv2d_small_fwd(imgd, dev_f32(wt, cuda_device), dev_f32(b, cuda_device), y, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), 1e-2, "conv1_1 fwd")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dw = torch.empty((3, 3, cin, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_small_wgrad(imgd, dev_bf16(dy, cuda_device), dw)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), 1e-4, "conv1_1 wgrad")


@pytest.mark.parametrize("tail_wide", [1, 0])
@pytest.mark.parametrize("shape", [(2, 5, 18, 4096, 2), (3, 5, 7, 1024, 4), (32, 5, 18, 4096, 2)])
def test_conv8_skinny_fwd_dgrad_wgrad(ops, cuda_device, shape, tail_wide):
    """conv8 (1x1, fc -> num_classes, FCN.py:86): the 16-byte-access forms for few pixels / wide channels
    (tail_wide = 1, the default) and the per-warp / per-element forms (0) against the oracle."""
    ops.ctx.set_tuning("tail_wide", tail_wide)
    try:
        _conv8_case(ops, cuda_device, shape)
    finally:
        ops.ctx.set_tuning("tail_wide", 1)


def _conv8_case(ops, cuda_device, shape):
    n, h, w, ci, co = shape
    rng = np.random.default_rng(1)
    x = bf16_grid(np.maximum(rng.standard_normal((n, h, w, ci)), 0))
    wt = (rng.standard_normal((1, 1, ci, co)) / 64).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    z = T.bias_add(T.conv2d_same(xt, wtt), torch.tensor(b))
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    xd = dev_bf16(x, cuda_device)
    ops.conv2d_small_fwd(xd, dev_f32(wt, cuda_device), dev_f32(b, cuda_device), y, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), T.relu(z).detach().numpy(), 1e-2, "conv8 fwd")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_small_dgrad(dev_bf16(dy, cuda_device), dev_f32(wt, cuda_device), dx, relu_mask=xd, scale=1.25)
    torch.cuda.synchronize()
    assert_close(host(dx), xt.grad.numpy() * (x > 0) * 1.25, 1e-2, "conv8 dgrad")
    dw = torch.empty((1, 1, ci, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_small_wgrad(xd, dev_bf16(dy, cuda_device), dw)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), 1e-4, "conv8 wgrad")


@pytest.mark.parametrize("tail_wide", [1, 0])
@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, s, f32 output/grad
    (2, 5, 18, 2, 512, 4, 2, False),      # conv_t1
    (32, 5, 18, 2, 512, 4, 2, False),     # conv_t1 at the benchmarked batch
    (3, 3, 5, 4, 64, 4, 2, False),        # four classes
    (1, 6, 8, 256, 2, 16, 8, True),       # conv_t3 (logits fp32)
    (2, 4, 4, 16, 24, 4, 2, False),
])
def test_deconv_small_fwd_dgrad_wgrad(ops, cuda_device, case, tail_wide):
    ops.ctx.set_tuning("tail_wide", tail_wide)
    try:
        _deconv_small_case(ops, cuda_device, case)
    finally:
        ops.ctx.set_tuning("tail_wide", 1)


def _deconv_small_case(ops, cuda_device, case):
    n, h, w, ci, co, k, s, f32 = case
    rng = np.random.default_rng(2)
    x = bf16_grid(np.maximum(rng.standard_normal((n, h, w, ci)), 0))
    wt = (rng.standard_normal((k, k, co, ci)) / np.sqrt(4 * ci)).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    res = None if f32 else bf16_grid(rng.standard_normal((n, h * s, w * s, co)))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y_ref = T.bias_add(T.conv2d_transpose_same(xt, wtt, (h * s, w * s), s), torch.tensor(b))
    if res is not None:
        y_ref = y_ref + torch.tensor(res)
    y = torch.empty((n, h * s, w * s, co), dtype=torch.float32 if f32 else torch.bfloat16, device=cuda_device)
    xd = dev_bf16(x, cuda_device)
    wd_ = dev_f32(wt, cuda_device)
    ops.deconv2d_small_fwd(xd, wd_, dev_f32(b, cuda_device), y, s,
                           residual=None if res is None else dev_bf16(res, cuda_device))
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), 1e-5 if f32 else 1e-2, f"deconv small fwd {case}")
    dy = rng.standard_normal((n, h * s, w * s, co)).astype(np.float32)
    if not f32:
        dy = bf16_grid(dy)
    y_ref.backward(torch.tensor(dy))
    dyd = dev_f32(dy, cuda_device) if f32 else dev_bf16(dy, cuda_device)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.deconv2d_small_dgrad(dyd, wd_, dx, s, relu_mask=xd)
    torch.cuda.synchronize()
    assert_close(host(dx), xt.grad.numpy() * (x > 0), 1e-2, f"deconv small dgrad {case}")
    dw = torch.empty((k, k, co, ci), dtype=torch.float32, device=cuda_device)
    ops.deconv2d_small_wgrad(xd, dyd, dw, s)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), 1e-4, f"deconv small wgrad {case}")


@pytest.mark.parametrize("cin", [3, 4])
def test_conv1_1_as_im2col_gemm(ops, cuda_device, cin):
    """conv1_1 routed through the 64-wide patch tensor + 1x1 tensor-core GEMM (fwd and wgrad)."""
    n, h, w, co = 2, 16, 24, 64
    rng = np.random.default_rng(10)
    img = rng.integers(0, 256, (n, h, w, cin), dtype=np.uint8)
    wt = bf16_grid(rng.standard_normal((3, 3, cin, co)) * 0.01)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    wtt = torch.tensor(wt, requires_grad=True)
    z = T.bias_add(T.conv2d_same(torch.tensor(img.astype(np.float32)), wtt), torch.tensor(b))
    P = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=cuda_device)
    ops.im2col_k64(torch.as_tensor(img).to(cuda_device), P, 3, 3)
    wk = ops.pack_im2col_weights(dev_f32(wt, cuda_device))
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(P, wk, dev_f32(b, cuda_device), y, 1, 1, relu=True)
    torch.cuda.synchronize()
    # patch tensor is exact: compare with a host im2col
    Pn = host(P)
    assert np.all(Pn[..., 9 * cin:] == 0)
    pad = np.pad(img.astype(np.float32), ((0, 0), (1, 1), (1, 1), (0, 0)))
    for t in range(9):
        ky, kx = divmod(t, 3)
        assert np.array_equal(Pn[..., t * cin:(t + 1) * cin], pad[:, ky:ky + h, kx:kx + w, :])
    assert_close(host(y), T.relu(z).detach().numpy(), 1e-2, "conv1_1 im2col fwd")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dw = torch.empty((1, 1, 64, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(P, dev_bf16(dy, cuda_device), dw, 1, 1)
    torch.cuda.synchronize()
    got = host(dw).reshape(64, co)
    assert np.all(got[9 * cin:] == 0)
    assert_close(got[:9 * cin].reshape(3, 3, cin, co), wtt.grad.numpy(), 2e-3, "conv1_1 im2col wgrad")


@pytest.mark.parametrize("case", [(2, 16, 24, 3, 64, "u8"), (3, 10, 37, 3, 64, "u8"), (1, 5, 7, 4, 64, "u8"),
                                  (2, 12, 20, 1, 128, "u8"), (2, 9, 33, 3, 256, "bf16"), (5, 4, 4, 3, 64, "u8")])
def test_conv1_1_fused_first_layer(ops, cuda_device, case):
    """conv1_1 with the 3x3 patches built in shared memory (no patch tensor): forward, and
    Conv2DBackpropFilter + BiasAddGrad in one pass over dy (FCN.py:52,340)."""
    n, h, w, cin, co, dt = case
    rng = np.random.default_rng(12)
    if dt == "u8":
        img = rng.integers(0, 256, (n, h, w, cin), dtype=np.uint8)
        xd = torch.as_tensor(img).to(cuda_device)
        xf = img.astype(np.float32)
    else:
        xf = bf16_grid(rng.standard_normal((n, h, w, cin)))
        xd = dev_bf16(xf, cuda_device)
    wt = bf16_grid(rng.standard_normal((3, 3, cin, co)) * 0.01)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    wtt = torch.tensor(wt, requires_grad=True)
    bt = torch.tensor(b, requires_grad=True)
    z = T.bias_add(T.conv2d_same(torch.tensor(xf), wtt), bt)
    wk = ops.pack_im2col_weights(dev_f32(wt, cuda_device))
    y = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_first_fwd(xd, wk, dev_f32(b, cuda_device), y, 3, 3, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), T.relu(z).detach().numpy(), 1e-2, f"first-layer fwd {case}")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dw = torch.full((3, 3, cin, co), float("nan"), dtype=torch.float32, device=cuda_device)
    db = torch.full((co,), float("nan"), dtype=torch.float32, device=cuda_device)
    ops.conv2d_first_wgrad(xd, dev_bf16(dy, cuda_device), dw, 3, 3, dbias=db)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), 2e-3, f"first-layer wgrad {case}")
    assert_close(host(db), bt.grad.numpy(), 2e-3, f"first-layer bias grad {case}")
    # and without the bias gradient
    dw2 = torch.empty_like(dw)
    ops.conv2d_first_wgrad(xd, dev_bf16(dy, cuda_device), dw2, 3, 3)
    torch.cuda.synchronize()
    assert torch.equal(dw, dw2)


def test_conv1_1_fused_matches_im2col_route_full_width(ops, cuda_device):
    """Same layer through both tensor-core routes at a KITTI-width image: identical operands, so the
    forward results agree to bf16 rounding of the same fp32 sums."""
    n, h, w, cin, co = 2, 32, 576, 3, 64
    rng = np.random.default_rng(13)
    img = torch.as_tensor(rng.integers(0, 256, (n, h, w, cin), dtype=np.uint8)).to(cuda_device)
    wt = dev_f32(bf16_grid(rng.standard_normal((3, 3, cin, co)) * 0.01), cuda_device)
    b = dev_f32((rng.standard_normal(co) * 0.1).astype(np.float32), cuda_device)
    wk = ops.pack_im2col_weights(wt)
    P = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=cuda_device)
    y1 = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    y2 = torch.empty_like(y1)
    ops.im2col_k64(img, P, 3, 3)
    ops.conv2d_fwd(P, wk, b, y1, 1, 1, relu=True)
    ops.conv2d_first_fwd(img, wk, b, y2, 3, 3, relu=True)
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, co))), cuda_device)
    dw1 = torch.empty((1, 1, 64, co), dtype=torch.float32, device=cuda_device)
    dw2 = torch.empty((3, 3, cin, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(P, dy, dw1, 1, 1)
    ops.conv2d_first_wgrad(img, dy, dw2, 3, 3)
    torch.cuda.synchronize()
    assert_close(host(y2), host(y1), 1e-2, "first-layer fwd vs im2col route")
    assert_close(host(dw2).reshape(27, co), host(dw1).reshape(64, co)[:27], 1e-3, "first-layer wgrad vs im2col route")


def test_conv1_1_fused_rejects_unsupported(ops, cuda_device):
    from semanticsegmentation_tensorflow_b200._lib import SegkError
    x = torch.zeros((1, 8, 8, 2), dtype=torch.uint8, device=cuda_device)
    wk = torch.zeros((1, 64, 64), dtype=torch.bfloat16, device=cuda_device)
    y = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(SegkError):
        ops.conv2d_first_fwd(x, wk, None, y, 3, 3)


def test_conv_t3_in_patch_space(ops, cuda_device):
    """conv_t3 (16x16 s8, Cout=2): fwd = 1x1 GEMM + col2im, bwd = patch gather + 1x1 GEMMs."""
    n, h, w, ci, co, k, s = 2, 5, 9, 256, 2, 16, 8
    rng = np.random.default_rng(11)
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = bf16_grid(rng.standard_normal((k, k, co, ci)) / np.sqrt(4 * ci))
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y_ref = T.bias_add(T.conv2d_transpose_same(xt, wtt, (h * s, w * s), s), torch.tensor(b))
    e = k * k * co
    wk, wd = ops.pack_matrix(dev_f32(wt, cuda_device).view(1, e, ci))
    xd = dev_bf16(x, cuda_device)
    yp = torch.empty((n, h, w, e), dtype=torch.float32, device=cuda_device)
    ops.conv2d_fwd(xd, wk, None, yp, 1, 1, relu=False)
    y = torch.empty((n, h * s, w * s, co), dtype=torch.float32, device=cuda_device)
    ops.deconv_col2im(yp, dev_f32(b, cuda_device), y, k, s)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), 1e-4, "conv_t3 patch fwd")
    dy = rng.standard_normal((n, h * s, w * s, co)).astype(np.float32)
    y_ref.backward(torch.tensor(dy))
    P = torch.empty((n, h, w, e), dtype=torch.bfloat16, device=cuda_device)
    ops.deconv_patch_gather(dev_f32(dy, cuda_device), P, k, s)
    dw = torch.empty((1, 1, e, ci), dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(P, xd, dw, 1, 1)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_dgrad(P, wd, dx, 1, 1)
    torch.cuda.synchronize()
    # dy is rounded to bf16 inside the patch tensor: 2^-9 relative per element
    assert_close(host(dw).reshape(k, k, co, ci), wtt.grad.numpy(), 5e-3, "conv_t3 patch wgrad")
    assert_close(host(dx), xt.grad.numpy(), 1e-2, "conv_t3 patch dgrad")


@pytest.mark.parametrize("co", [2, 4])
def test_full_resolution_1x1_head(ops, cuda_device, co):
    """64 -> num_classes 1x1 `final_conv` head (FCDenseNet.py:157) on >= 65536 pixels: the streaming
    thread-per-pixel / two-stage-reduction kernels, fp32 logits out."""
    n, h, w, ci = 2, 160, 288, 64
    rng = np.random.default_rng(30)
    x = bf16_grid(np.maximum(rng.standard_normal((n, h, w, ci)), 0))
    wt = (rng.standard_normal((1, 1, ci, co)) / 8).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    z = T.conv2d_same(xt, wtt)
    xd = dev_bf16(x, cuda_device)
    y = torch.empty((n, h, w, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_small_fwd(xd, dev_f32(wt, cuda_device), None, y, relu=False)
    torch.cuda.synchronize()
    assert_close(host(y), z.detach().numpy(), 1e-5, "head fwd")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_small_dgrad(dev_bf16(dy, cuda_device), dev_f32(wt, cuda_device), dx, relu_mask=xd)
    dw = torch.empty((1, 1, ci, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_small_wgrad(xd, dev_bf16(dy, cuda_device), dw)
    torch.cuda.synchronize()
    assert_close(host(dx), xt.grad.numpy() * (x > 0), 1e-2, "head dgrad")
    assert_close(host(dw), wtt.grad.numpy(), 1e-4, "head wgrad")


@pytest.mark.parametrize("case", [(2, 16, 24, 64, 2, 3), (1, 9, 13, 32, 4, 3), (1, 12, 10, 64, 2, 5)])
def test_skinny_kxk_head(ops, cuda_device, case):
    """3x3 (5x5) conv to num_classes on CUDA cores: SegNet's conv26 (SegNet.py:80), fwd / dgrad / wgrad."""
    n, h, w, ci, co, k = case
    rng = np.random.default_rng(60)
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = (rng.standard_normal((k, k, ci, co)) / np.sqrt(k * k * ci)).astype(np.float32)
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    xt, wtt = torch.tensor(x, requires_grad=True), torch.tensor(wt, requires_grad=True)
    z = T.bias_add(T.conv2d_same(xt, wtt), torch.tensor(b))
    y = torch.empty((n, h, w, co), dtype=torch.float32, device=cuda_device)
    xd, wd_ = dev_bf16(x, cuda_device), dev_f32(wt, cuda_device)
    ops.conv2d_small_fwd(xd, wd_, dev_f32(b, cuda_device), y, relu=False)
    torch.cuda.synchronize()
    assert_close(host(y), z.detach().numpy(), 1e-5, f"skinny kxk fwd {case}")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dyd = dev_bf16(dy, cuda_device)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_small_dgrad(dyd, wd_, dx, relu_mask=xd)
    torch.cuda.synchronize()
    assert_close(host(dx), xt.grad.numpy() * (x > 0), 1e-2, f"skinny kxk dgrad {case}")
    dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.conv2d_small_wgrad(xd, dyd, dw)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), 1e-4, f"skinny kxk wgrad {case}")


@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, s
    (2, 5, 9, 256, 2, 16, 8),        # conv_t3 (FCN.py:98-107)
    (32, 20, 72, 256, 2, 16, 8),     # ... at the benchmarked shape (only spot-checked against the oracle below)
    (1, 7, 10, 64, 4, 16, 8),        # four classes: R = 256
    (3, 6, 5, 128, 4, 8, 4),         # k = 8, s = 4: R = 64
])
def test_conv_t3_phase_packed(ops, cuda_device, case):
    """The phase-packed form of the tiny-Cout transposed conv: one 4-tap GEMM over the block grid with the logits
    written by the epilogue; backward from the re-blocked dy.  Against oracle/tf_ops.conv2d_transpose_same."""
    n, h, w, ci, co, k, s = case
    rng = np.random.default_rng(12)
    full = n * h * w <= 4000
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = bf16_grid(rng.standard_normal((k, k, co, ci)) / np.sqrt(4 * ci))
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    dy = (rng.standard_normal((n, h * s, w * s, co)) / 64).astype(np.float32)
    bf, bt = ops.pack_deconv_packed(dev_f32(wt, cuda_device), s)
    xd = dev_bf16(x, cuda_device)
    y = torch.full((n, h * s, w * s, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.deconv2d_packed_fwd(xd, bf, dev_f32(b, cuda_device), y, k, s)
    r = s * s * co
    dyb = torch.full((n, h + 1, w + 1, r), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.deconv_pack_dy(dev_f32(dy, cuda_device), dyb, s)
    dx = torch.full((n, h, w, ci), 7.0, dtype=torch.bfloat16, device=cuda_device)
    cs = torch.full((ci,), 7.0, dtype=torch.float32, device=cuda_device)
    ops.deconv2d_packed_dgrad(dyb, bt, dx, co, k, s, colsum=cs)
    dw = torch.full((k, k, co, ci), 7.0, dtype=torch.float32, device=cuda_device)
    dwt = torch.empty((4, ci, r), dtype=torch.float32, device=cuda_device)
    ops.deconv2d_packed_wgrad(xd, dyb, dw, dwt, k, s)
    torch.cuda.synchronize()
    sl = slice(None) if full else slice(0, 2)          # the oracle on the first two images of the large case
    xt = torch.tensor(x[sl], requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y_ref = T.bias_add(T.conv2d_transpose_same(xt, wtt, (h * s, w * s), s), torch.tensor(b))
    assert_close(host(y)[sl], y_ref.detach().numpy(), 1e-4, f"packed deconv fwd {case}")
    y_ref.backward(torch.tensor(bf16_grid(dy[sl])))     # dy is rounded to bf16 in the block tensor
    assert_close(host(dx)[sl], xt.grad.numpy(), 1e-2, f"packed deconv dgrad {case}")
    if full:
        assert_close(host(dw), wtt.grad.numpy(), 2e-3, f"packed deconv wgrad {case}")
        gx = xt.grad.numpy()
        np.testing.assert_allclose(host(cs), gx.astype(np.float64).sum(axis=(0, 1, 2)), rtol=0,
                                   atol=4e-3 * np.abs(gx).max() * np.sqrt(n * h * w))
    else:
        # wgrad is additive over images: the full batch equals the sum over two halves
        dwa = torch.empty_like(dw)
        dwb = torch.empty_like(dw)
        half = n // 2
        for lo, hi, out in ((0, half, dwa), (half, n, dwb)):
            d2 = torch.empty((hi - lo, h + 1, w + 1, r), dtype=torch.bfloat16, device=cuda_device)
            ops.deconv_pack_dy(dev_f32(dy[lo:hi], cuda_device), d2, s)
            ops.deconv2d_packed_wgrad(xd[lo:hi].contiguous(), d2, out, dwt, k, s)
        torch.cuda.synchronize()
        assert_close(host(dw), host(dwa) + host(dwb), 1e-4, f"packed deconv wgrad additivity {case}")
