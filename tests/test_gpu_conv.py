"""GPU parity: tensor-core conv family (through the C ABI) vs the CPU oracle (oracle/tf_ops.py).

Inputs are drawn on the bf16 grid so the only differences are accumulation order (fp32 in
both) and the final bf16 store: tolerance 1e-2 of the tensor max for bf16 outputs
(bf16 ulp = 2^-8 relative), 2e-3 for fp32 outputs (north_star: rtol 2e-2 under bf16)."""
import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from tests.gpu_util import assert_close, bf16_grid, dev_bf16, dev_f32, host

pytestmark = pytest.mark.gpu

TOL_BF16 = 1e-2
TOL_F32 = 2e-3


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


CONV_SHAPES = [
    # N, H, W, Cin, Cout, k
    (2, 16, 24, 64, 64, 3),
    (1, 8, 8, 128, 256, 3),
    (3, 10, 12, 128, 64, 1),
    (2, 5, 18, 64, 128, 7),
    (1, 20, 72, 256, 512, 3),
    (4, 10, 36, 512, 512, 3),
    (2, 40, 144, 128, 256, 3),
    (1, 7, 9, 64, 192, 3),       # ragged box, Cout = 3 x 64
    (1, 5, 18, 256, 128, 7),     # few tiles, long K walk -> split-K path (fwd and dgrad)
    (4, 48, 100, 128, 256, 3),   # 150 tiles on 148 SMs: a second, nearly empty wave
]


def _conv_case(shape, seed=0):
    n, h, w, ci, co, k = shape
    rng = np.random.default_rng(seed)
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = bf16_grid(rng.standard_normal((k, k, ci, co)) / np.sqrt(k * k * ci))
    b = rng.standard_normal(co).astype(np.float32) * 0.1
    return x, wt, b


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv2d_fwd_bias_relu(ops, cuda_device, shape):
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape)
    ref = T.relu(T.bias_add(T.conv2d_same(torch.tensor(x), torch.tensor(wt)), torch.tensor(b))).numpy()
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(dev_bf16(x, cuda_device), wk, dev_f32(b, cuda_device), y, k, k, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), ref, TOL_BF16, f"conv fwd {shape}")


def test_conv2d_fwd_residual_f32_no_relu(ops, cuda_device):
    shape = (2, 12, 20, 128, 128, 3)
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 1)
    res = bf16_grid(np.random.default_rng(2).standard_normal((n, h, w, co)))
    ref = (T.bias_add(T.conv2d_same(torch.tensor(x), torch.tensor(wt)), torch.tensor(b)) + torch.tensor(res)).numpy()
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    y = torch.empty((n, h, w, co), dtype=torch.float32, device=cuda_device)
    ops.conv2d_fwd(dev_bf16(x, cuda_device), wk, dev_f32(b, cuda_device), y, k, k, relu=False,
                   residual=dev_bf16(res, cuda_device))
    torch.cuda.synchronize()
    assert_close(host(y), ref, TOL_F32, "conv fwd residual f32")


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv2d_dgrad(ops, cuda_device, shape):
    n, h, w, ci, co, k = shape
    x, wt, _ = _conv_case(shape, 3)
    rng = np.random.default_rng(4)
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    act = bf16_grid(rng.standard_normal((n, h, w, ci)))          # forward activation -> ReLU mask
    res = bf16_grid(rng.standard_normal((n, h, w, ci)))
    xt = torch.tensor(x, requires_grad=True)
    T.conv2d_same(xt, torch.tensor(wt)).backward(torch.tensor(dy))
    ref = ((xt.grad.numpy() + res) * (act > 0)) * 1.25
    _, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_dgrad(dev_bf16(dy, cuda_device), wd, dx, k, k, relu_mask=dev_bf16(act, cuda_device),
                     residual=dev_bf16(res, cuda_device), scale=1.25)
    torch.cuda.synchronize()
    assert_close(host(dx), ref, TOL_BF16, f"conv dgrad {shape}")
    # the same call with the fused BiasAddGrad of the producer layer: column sums of dx (fp32 epilogue values)
    if ci % 32 == 0:
        dx2 = torch.full((n, h, w, ci), 7.0, dtype=torch.bfloat16, device=cuda_device)
        cs = torch.full((ci,), 7.0, dtype=torch.float32, device=cuda_device)
        ops.conv2d_dgrad(dev_bf16(dy, cuda_device), wd, dx2, k, k, relu_mask=dev_bf16(act, cuda_device),
                         residual=dev_bf16(res, cuda_device), scale=1.25, colsum=cs)
        torch.cuda.synchronize()
        assert torch.equal(dx2, dx)
        ref_cs = ref.astype(np.float64).sum(axis=(0, 1, 2))
        np.testing.assert_allclose(host(cs), ref_cs, rtol=0, atol=4e-3 * np.abs(ref).max() * np.sqrt(n * h * w))


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv2d_wgrad(ops, cuda_device, shape):
    n, h, w, ci, co, k = shape
    if n * h * w < 64:
        pytest.skip("needs >= 64 pixels")
    x, wt, _ = _conv_case(shape, 5)
    dy = bf16_grid(np.random.default_rng(6).standard_normal((n, h, w, co)))
    wtt = torch.tensor(wt, requires_grad=True)
    T.conv2d_same(torch.tensor(x), wtt).backward(torch.tensor(dy))
    ref = wtt.grad.numpy()
    dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)   # must be overwritten
    ops.conv2d_wgrad(dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device), dw, k, k)
    torch.cuda.synchronize()
    assert_close(host(dw), ref, TOL_F32, f"conv wgrad {shape}")
    # accumulate=1 adds on top
    ops.conv2d_wgrad(dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device), dw, k, k, accumulate=True)
    torch.cuda.synchronize()
    assert_close(host(dw), 2 * ref, TOL_F32, f"conv wgrad accumulate {shape}")


DECONV_SHAPES = [
    # N, H, W, Cin, Cout  (k=4, s=2)
    (2, 5, 9, 64, 64),
    (2, 10, 36, 512, 256),
    (1, 8, 8, 128, 192),
]


@pytest.mark.parametrize("shape", DECONV_SHAPES)
def test_deconv2d_tc_fwd_dgrad_wgrad(ops, cuda_device, shape):
    n, h, w, ci, co = shape
    k, s = 4, 2
    rng = np.random.default_rng(7)
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = bf16_grid(rng.standard_normal((k, k, co, ci)) / np.sqrt(4 * ci))
    b = rng.standard_normal(co).astype(np.float32) * 0.1
    res = bf16_grid(rng.standard_normal((n, h * s, w * s, co)))
    dy = bf16_grid(rng.standard_normal((n, h * s, w * s, co)))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y_ref = T.bias_add(T.conv2d_transpose_same(xt, wtt, (h * s, w * s), s), torch.tensor(b)) + torch.tensor(res)
    y_ref.backward(torch.tensor(dy))
    wk, wd = ops.pack_deconv_weights(dev_f32(wt, cuda_device), s)
    y = torch.empty((n, h * s, w * s, co), dtype=torch.bfloat16, device=cuda_device)
    ops.deconv2d_fwd(dev_bf16(x, cuda_device), wk, dev_f32(b, cuda_device), y, k, s, residual=dev_bf16(res, cuda_device))
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), TOL_BF16, f"deconv fwd {shape}")
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    cs = torch.full((ci,), 7.0, dtype=torch.float32, device=cuda_device)
    ops.deconv2d_dgrad(dev_bf16(dy, cuda_device), wd, dx, k, s, colsum=cs)
    torch.cuda.synchronize()
    assert_close(host(dx), xt.grad.numpy(), TOL_BF16, f"deconv dgrad {shape}")
    gx = xt.grad.numpy()
    np.testing.assert_allclose(host(cs), gx.astype(np.float64).sum(axis=(0, 1, 2)), rtol=0,
                               atol=4e-3 * np.abs(gx).max() * np.sqrt(n * h * w))
    if n * h * w >= 64:
        dw = torch.empty((k, k, co, ci), dtype=torch.float32, device=cuda_device)
        ops.deconv2d_wgrad(dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device), dw, k, s)
        torch.cuda.synchronize()
        assert_close(host(dw), wtt.grad.numpy(), TOL_F32, f"deconv wgrad {shape}")


def test_unsupported_shape_is_an_error_not_a_fallback(ops, cuda_device):
    from semanticsegmentation_tensorflow_b200._lib import SegkError, SEGK_EINVAL
    x = torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device=cuda_device)
    wk = torch.zeros((9, 64, 48), dtype=torch.bfloat16, device=cuda_device)
    y = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(SegkError) as ei:
        ops.conv2d_fwd(x, wk, None, y, 3, 3)
    assert ei.value.code == SEGK_EINVAL and "no fallback" in str(ei.value)


WSLAB_SHAPES = [
    # N, H, W, Cin, Cout: 3x3 layers routed to the slab-formulated wgrad (wslab = 2 forces it on small maps)
    (2, 16, 64, 64, 64),          # all nine taps per item, tap pairs out of one slab
    (1, 10, 37, 64, 128),         # ragged rows / columns, two 64-channel output tiles
    (2, 12, 40, 128, 128),        # one filter row per item, (tap, both channel chunks) pairs, N = 128
    (1, 9, 70, 128, 64),          # ... N = 64
    (1, 64, 96, 64, 64),          # large enough for the automatic route
    (3, 5, 33, 128, 256),
]


@pytest.mark.parametrize("shape", WSLAB_SHAPES)
def test_slab_wgrad(ops, cuda_device, shape):
    n, h, w, ci, co = shape
    k = 3
    x, wt, _ = _conv_case((n, h, w, ci, co, k), 30)
    dy = bf16_grid(np.random.default_rng(31).standard_normal((n, h, w, co)))
    wtt = torch.tensor(wt, requires_grad=True)
    T.conv2d_same(torch.tensor(x), wtt).backward(torch.tensor(dy))
    ref = wtt.grad.numpy()
    xd, dyd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device)
    got = {}
    try:
        for mode in (2, 0):
            ops.ctx.set_tuning("wslab", mode)
            dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)   # must be overwritten
            ops.conv2d_wgrad(xd, dyd, dw, k, k)
            torch.cuda.synchronize()
            got[mode] = host(dw)
            if mode == 2:
                ops.conv2d_wgrad(xd, dyd, dw, k, k, accumulate=True)
                torch.cuda.synchronize()
                assert_close(host(dw), 2 * ref, TOL_F32, f"slab wgrad accumulate {shape}")
                # deterministic: per-split partial sums + ordered reduction, no atomics
                dw2 = torch.empty_like(dw)
                ops.conv2d_wgrad(xd, dyd, dw2, k, k)
                torch.cuda.synchronize()
                assert np.array_equal(host(dw2), got[2])
    finally:
        ops.ctx.set_tuning("wslab", 1)
    assert_close(got[2], ref, TOL_F32, f"slab wgrad {shape}")
    assert_close(got[2], got[0], TOL_F32, f"slab wgrad vs tap-wise wgrad {shape}")


def test_slab3_forced_on_dgrad_with_mask(ops, cuda_device):
    """slab3 = 2 also sends the masked / scaled dgrad epilogue through the kx-fused kernel."""
    n, h, w, ci, co, k = 1, 64, 96, 64, 64, 3
    x, wt, _ = _conv_case((n, h, w, ci, co, k), 24)
    rng = np.random.default_rng(25)
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    xt = torch.tensor(x, requires_grad=True)
    T.conv2d_same(xt, torch.tensor(wt)).backward(torch.tensor(dy))
    _, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.ctx.set_tuning("slab3", 2)
    try:
        ops.conv2d_dgrad(dev_bf16(dy, cuda_device), wd, dx, k, k, relu_mask=dev_bf16(x, cuda_device), scale=0.5)
        torch.cuda.synchronize()
    finally:
        ops.ctx.set_tuning("slab3", 1)
    assert_close(host(dx), xt.grad.numpy() * (x > 0) * 0.5, TOL_BF16, "slab3 dgrad with mask")


SLAB_SHAPES = [
    # N, H, W, Cin, Cout  (3x3; routed to the haloed-slab kernel: Cout <= 128 and H*W >= 4096)
    (2, 64, 96, 64, 64),
    (1, 80, 288, 64, 128),
    (1, 160, 576, 64, 64),
    (2, 40, 144, 128, 128),
    (1, 64, 70, 128, 64),        # W not a multiple of 30: ragged last column tile
    (1, 64, 70, 64, 64),         # the same through the kx-fused N = 192 kernel (Cin = 64)
    (3, 8, 576, 64, 128),        # two channel tiles per CTA grid, short images
]


@pytest.mark.parametrize("shape", SLAB_SHAPES)
def test_slab_conv_fwd_and_dgrad(ops, cuda_device, shape):
    n, h, w, ci, co = shape
    k = 3
    x, wt, b = _conv_case((n, h, w, ci, co, k), 20)
    rng = np.random.default_rng(21)
    res = bf16_grid(rng.standard_normal((n, h, w, co)))
    xt = torch.tensor(x, requires_grad=True)
    z = T.bias_add(T.conv2d_same(xt, torch.tensor(wt)), torch.tensor(b)) + torch.tensor(res)
    ref = T.relu(z).detach().numpy()
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(dev_bf16(x, cuda_device), wk, dev_f32(b, cuda_device), y, k, k, relu=True,
                   residual=dev_bf16(res, cuda_device))
    torch.cuda.synchronize()
    assert_close(host(y), ref, TOL_BF16, f"slab fwd {shape}")
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    z.backward(torch.tensor(dy))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    cs = torch.full((ci,), 7.0, dtype=torch.float32, device=cuda_device)
    ops.conv2d_dgrad(dev_bf16(dy, cuda_device), wd, dx, k, k, relu_mask=dev_bf16(x, cuda_device), scale=0.5, colsum=cs)
    torch.cuda.synchronize()
    ref_dx = xt.grad.numpy() * (x > 0) * 0.5
    assert_close(host(dx), ref_dx, TOL_BF16, f"slab dgrad {shape}")
    np.testing.assert_allclose(host(cs), ref_dx.astype(np.float64).sum(axis=(0, 1, 2)), rtol=0,
                               atol=4e-3 * np.abs(ref_dx).max() * np.sqrt(n * h * w))


def test_slab_forced_for_wide_layers(ops, cuda_device):
    """slab = 2 forces the slab kernel wherever it is legal (BLOCK_N = 256 instance)."""
    ops.ctx.set_tuning("slab", 2)
    try:
        _slab_forced_body(ops, cuda_device)
    finally:
        ops.ctx.set_tuning("slab", 1)


def test_slab3_matches_plain_slab_and_handles_three_channel_tiles(ops, cuda_device):
    """Cin = 64 layers take the kx-fused kernel (three taps per N = 192 MMA, weights resident);
    slab3 = 0 sends the same call through the tap-wise slab kernel: same sums, fp32 reassociated."""
    n, h, w, ci, co, k = 2, 32, 150, 64, 192, 3
    x, wt, b = _conv_case((n, h, w, ci, co, k), 23)
    ref = T.relu(T.bias_add(T.conv2d_same(torch.tensor(x), torch.tensor(wt)), torch.tensor(b))).numpy()
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd = dev_bf16(x, cuda_device), dev_f32(b, cuda_device)
    ys = []
    ops.ctx.set_tuning("slab", 2)
    try:
        for fused in (2, 0):
            ops.ctx.set_tuning("slab3", fused)
            y = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=cuda_device)
            ops.conv2d_fwd(xd, wk, bd, y, k, k, relu=True)
            torch.cuda.synchronize()
            ys.append(host(y))
    finally:
        ops.ctx.set_tuning("slab", 1)
        ops.ctx.set_tuning("slab3", 1)
    assert_close(ys[0], ref, TOL_BF16, "slab3 fwd, 3 channel tiles")
    assert_close(ys[0], ys[1], TOL_BF16, "slab3 vs tap-wise slab")


def _slab_forced_body(ops, cuda_device):
    shape = (1, 40, 144, 128, 256, 3)
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 22)
    ref = T.relu(T.bias_add(T.conv2d_same(torch.tensor(x), torch.tensor(wt)), torch.tensor(b))).numpy()
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(dev_bf16(x, cuda_device), wk, dev_f32(b, cuda_device), y, k, k, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), ref, TOL_BF16, "slab forced BN=256")


def _guarded(shape, dtype, device, pad=4096):
    """A tensor view in the middle of a larger byte buffer filled with 0xA5: (view, check) where check()
    asserts that the guard bytes before and after the view are untouched."""
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    raw = torch.full((pad + n + pad,), 0xA5, dtype=torch.uint8, device=device)
    view = raw[pad:pad + n].view(dtype).view(shape)

    def check(what):
        torch.cuda.synchronize()
        assert bool((raw[:pad] == 0xA5).all()) and bool((raw[pad + n:] == 0xA5).all()), f"{what}: wrote outside its output"

    return view, check


@pytest.mark.parametrize("shape", [(3, 10, 37, 64, 64), (2, 12, 70, 64, 128), (1, 9, 45, 128, 128), (2, 8, 33, 64, 64)])
def test_slab_kernels_stay_inside_their_outputs(ops, cuda_device, shape):
    """Ragged maps through the TMA-store epilogues (slab / slab3), the slab wgrad and the fused first
    layer, with guard bytes around every output (compute-sanitizer is not available on the GPU pool)."""
    n, h, w, ci, co = shape
    k = 3
    x, wt, b = _conv_case((n, h, w, ci, co, k), 50)
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd = dev_bf16(x, cuda_device), dev_f32(b, cuda_device)
    dy = dev_bf16(bf16_grid(np.random.default_rng(51).standard_normal((n, h, w, co))), cuda_device)
    try:
        ops.ctx.set_tuning("slab", 2)
        ops.ctx.set_tuning("wslab", 2)
        for s3 in (2, 0):
            ops.ctx.set_tuning("slab3", s3)
            y, chk = _guarded((n, h, w, co), torch.bfloat16, cuda_device)
            ops.conv2d_fwd(xd, wk, bd, y, k, k, relu=True)          # (H % 4 != 0 goes through the igemm kernel)
            chk(f"conv fwd slab3={s3}")
            dx, chk = _guarded((n, h, w, ci), torch.bfloat16, cuda_device)
            ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask=xd)
            chk(f"conv dgrad slab3={s3}")
        dw, chk = _guarded((k, k, ci, co), torch.float32, cuda_device)
        ops.conv2d_wgrad(xd, dy, dw, k, k)
        chk("slab wgrad")
    finally:
        ops.ctx.set_tuning("slab", 1)
        ops.ctx.set_tuning("slab3", 1)
        ops.ctx.set_tuning("wslab", 1)
    img = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=cuda_device)
    wk1 = ops.pack_im2col_weights(dev_f32(_conv_case((1, 4, 4, 3, 64, 3), 52)[1], cuda_device))
    y1, chk = _guarded((n, h, w, 64), torch.bfloat16, cuda_device)
    ops.conv2d_first_fwd(img, wk1, None, y1, 3, 3, relu=True)
    chk("first-layer fwd")
    dw1, chk1 = _guarded((3, 3, 3, 64), torch.float32, cuda_device)
    db1, chk2 = _guarded((64,), torch.float32, cuda_device)
    dy1 = dev_bf16(bf16_grid(np.random.default_rng(53).standard_normal((n, h, w, 64))), cuda_device)
    ops.conv2d_first_wgrad(img, dy1, dw1, 3, 3, dbias=db1)
    chk1("first-layer wgrad")
    chk2("first-layer bias grad")


@pytest.mark.parametrize("shape", [(16, 5, 18, 256, 256, 7), (2, 5, 18, 512, 1024, 7), (9, 5, 18, 128, 512, 5)])
def test_team_stream_k_fwd_dgrad_and_wgrad_box_skipping(ops, cuda_device, shape):
    """conv6-like layers (k > map height, few output tiles, long K walk): the lockstep tap-split schedule of
    igemm_kernel (a CTA owns (tile, tap) pairs of up to two pixel tiles, tiles with different tap counts, pieces
    of a tile in different partial slices) against the oracle and against the plain split-K path; and the weight
    gradient with all-padding (pixel box, tap) pairs skipped."""
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 40)
    rng = np.random.default_rng(41)
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    act = bf16_grid(rng.standard_normal((n, h, w, ci)))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    z = T.conv2d_same(xt, wtt)
    y_ref = T.relu(T.bias_add(z, torch.tensor(b))).detach().numpy()
    z.backward(torch.tensor(dy))
    dx_ref = xt.grad.numpy() * (act > 0)
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, dyd, actd, bd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device), dev_bf16(act, cuda_device), dev_f32(b, cuda_device)
    out = {}
    try:
        for mode in (1, 0):
            ops.ctx.set_tuning("teamk", mode)
            y = torch.full((n, h, w, co), 7.0, dtype=torch.bfloat16, device=cuda_device)
            dx = torch.full((n, h, w, ci), 7.0, dtype=torch.bfloat16, device=cuda_device)
            ops.conv2d_fwd(xd, wk, bd, y, k, k, relu=True)
            ops.conv2d_dgrad(dyd, wd, dx, k, k, relu_mask=actd)
            torch.cuda.synchronize()
            assert_close(host(y), y_ref, TOL_BF16, f"team-K fwd {shape} mode {mode}")
            assert_close(host(dx), dx_ref, TOL_BF16, f"team-K dgrad {shape} mode {mode}")
            out[mode] = (host(y), host(dx))
            # deterministic: ordered partial sums, no atomics
            y2 = torch.empty_like(y)
            ops.conv2d_fwd(xd, wk, bd, y2, k, k, relu=True)
            torch.cuda.synchronize()
            assert torch.equal(y, y2)
    finally:
        ops.ctx.set_tuning("teamk", 1)
    dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(xd, dyd, dw, k, k)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), TOL_F32, f"wgrad with box skipping {shape}")


@pytest.mark.parametrize("shape", [(4, 48, 100, 128, 256, 3), (32, 10, 36, 128, 512, 3), (8, 16, 16, 4096, 2560, 1)])
def test_hybrid_schedule_matches_plain_waves(ops, cuda_device, shape):
    """A partly filled last wave (150 / 180 tiles on 148 SMs): whole tiles leave through the TMA-store epilogue, the
    remainder tiles as K-split fp32 partial sums finished by epilogue_finish_tiles_kernel (channel-tile-fastest and
    pixel-tile-fastest orders).  Against the oracle, against the plain schedule, and deterministic."""
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 60)
    res = bf16_grid(np.random.default_rng(61).standard_normal((n, h, w, co)))
    ref = T.relu(T.bias_add(T.conv2d_same(torch.tensor(x), torch.tensor(wt)), torch.tensor(b)) + torch.tensor(res)).numpy()
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd, rd = dev_bf16(x, cuda_device), dev_f32(b, cuda_device), dev_bf16(res, cuda_device)
    out = {}
    try:
        for mode in (1, 0):
            ops.ctx.set_tuning("hybrid", mode)
            y = torch.full((n, h, w, co), 7.0, dtype=torch.bfloat16, device=cuda_device)
            n0 = ops.ctx.launches
            ops.conv2d_fwd(xd, wk, bd, y, k, k, relu=True, residual=rd)
            torch.cuda.synchronize()
            out[mode] = (y, ops.ctx.launches - n0)
            assert_close(host(y), ref, TOL_BF16, f"hybrid={mode} fwd {shape}")
        assert out[1][1] == 2 and out[0][1] == 1, (out[1][1], out[0][1])       # igemm + finish vs igemm alone
        y2 = torch.empty_like(out[1][0])
        ops.ctx.set_tuning("hybrid", 1)
        ops.conv2d_fwd(xd, wk, bd, y2, k, k, relu=True, residual=rd)
        torch.cuda.synchronize()
        assert torch.equal(y2, out[1][0])
        # fp32 output through the same path
        yf = torch.full((n, h, w, co), 7.0, dtype=torch.float32, device=cuda_device)
        ops.conv2d_fwd(xd, wk, bd, yf, k, k, relu=True, residual=rd)
        torch.cuda.synchronize()
        assert_close(host(yf), ref, TOL_F32, f"hybrid fwd f32 {shape}")
    finally:
        ops.ctx.set_tuning("hybrid", 0)


@pytest.mark.parametrize("shape", [
    (2, 16, 60, 64, 64, 3),      # slab3 (4 x 30 tiles, Cin = Cout = 64)
    (2, 64, 96, 64, 128, 3),     # slab_kernel<128>
    (3, 20, 72, 128, 256, 3),    # igemm<256>, even box
    (2, 10, 36, 256, 512, 3),    # igemm<256>, two channel tiles, few tiles (no split-K when the pool is fused)
    (1, 8, 62, 64, 64, 3),       # ragged right edge (62 = 2 x 30 + 2)
    (2, 6, 10, 64, 192, 1),      # 1x1, Cout = 3 x 64
])
@pytest.mark.parametrize("relu", [True, False])
def test_conv2d_fwd_pool_bit_exact_vs_two_kernels(ops, cuda_device, shape, relu):
    """conv -> ReLU -> max_pool 2x2 (FCN.py:54-76) with the pool in the conv epilogue: pooled values and first-max
    indices bit-identical to segk_conv2d_fwd + segk_maxpool2x2_fwd (tie-heavy: ReLU zeros), y identical when stored,
    untouched with pool_only."""
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 70)
    wk, _ = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd = dev_bf16(x, cuda_device), dev_f32(b - 0.3, cuda_device)          # shifted bias: many ReLU zeros -> ties
    y0 = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(xd, wk, bd, y0, k, k, relu=relu)
    # fused, pre-pool tensor stored too: the pool of exactly that tensor (the plain conv may take a split-K schedule on
    # few-tile shapes, i.e. another fp32 summation order: y is compared to tolerance, the pool bit for bit)
    y1 = torch.full_like(y0, 7.0)
    p1 = torch.full((n, h // 2, w // 2, co), 7.0, dtype=torch.bfloat16, device=cuda_device)
    i1 = torch.full((n, h // 2, w // 2, co), 9, dtype=torch.uint8, device=cuda_device)
    ops.conv2d_fwd_pool(xd, wk, bd, y1, p1, i1, k, k, relu=relu, pool_only=False)
    p0, i0 = torch.empty_like(p1), torch.empty_like(i1)
    ops.maxpool_fwd(y1, p0, i0)
    torch.cuda.synchronize()
    assert_close(host(y1), host(y0), TOL_BF16, f"fused-pool conv output {shape}")
    assert torch.equal(p1, p0), f"pooled values {shape}"
    assert torch.equal(i1, i0), f"pool indices {shape}"
    # pool only: same pooled tensor and indices, y untouched
    y2 = torch.full_like(y0, 7.0)
    p2, i2 = torch.full_like(p1, 7.0), torch.full_like(i1, 9)
    ops.conv2d_fwd_pool(xd, wk, bd, y2, p2, i2, k, k, relu=relu, pool_only=True)
    torch.cuda.synchronize()
    assert torch.equal(p2, p1) and torch.equal(i2, i1), f"pool_only {shape}"
    assert bool((y2 == 7.0).all()), "pool_only must not write the pre-pool tensor"
    if relu:
        assert float((p0 == 0).float().mean()) > 0.05          # the tie case is exercised (SIMD integer path)


def _pack_bits(y):
    """[N,H,W,C] bf16 tensor -> int32 [N,H,W,C/32] words, bit i of word w = (y[..., 32 w + i] > 0)."""
    n, h, w, c = y.shape
    b = (y.float() > 0).to(torch.int64).view(n, h, w, c // 32, 32)
    weights = (1 << torch.arange(32, device=y.device, dtype=torch.int64))
    words = (b * weights).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


@pytest.mark.parametrize("shape", [
    (2, 16, 60, 64, 64, 3),      # slab3
    (2, 64, 96, 64, 128, 3),     # slab_kernel<128>
    (3, 20, 72, 128, 256, 3),    # igemm<256>
    (1, 5, 18, 256, 128, 7),     # few tiles, long K walk: split / tap-split schedules (finish kernels)
    (2, 6, 10, 64, 192, 1),      # 1x1, three channel tiles
])
def test_relu_mask_bits_fwd_and_dgrad(ops, cuda_device, shape):
    """1-bit ReLU masks: the forward epilogue's relu_bits equal [stored y > 0] exactly, and a dgrad that takes the
    producer's mask as bits gives the same bits as one that reads the bf16 activation."""
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 80)
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd = dev_bf16(x, cuda_device), dev_f32(b - 0.2, cuda_device)
    y0 = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    y1 = torch.empty_like(y0)
    bits = torch.full((n, h, w, co // 32), 0x5a5a5a5a, dtype=torch.int32, device=cuda_device)
    ops.conv2d_fwd(xd, wk, bd, y0, k, k, relu=True)
    ops.conv2d_fwd(xd, wk, bd, y1, k, k, relu=True, relu_bits=bits)
    torch.cuda.synchronize()
    assert torch.equal(y1, y0)
    assert torch.equal(bits, _pack_bits(y1)), f"relu bits {shape}"
    assert 0.05 < float((y1 > 0).float().mean()) < 0.95
    # consumer side: dgrad of a layer whose INPUT is `act` (mask of shape dx)
    rng = np.random.default_rng(81)
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, co))), cuda_device)
    act = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    res = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    abits = _pack_bits(act)
    dx0 = torch.full((n, h, w, ci), 7.0, dtype=torch.bfloat16, device=cuda_device)
    dx1 = torch.full_like(dx0, 7.0)
    ops.conv2d_dgrad(dy, wd, dx0, k, k, relu_mask=act, residual=res, scale=1.25)
    ops.conv2d_dgrad(dy, wd, dx1, k, k, relu_mask_bits=abits, residual=res, scale=1.25)
    torch.cuda.synchronize()
    assert torch.equal(dx1, dx0), f"dgrad with mask bits {shape}"
    if ci % 32 == 0:
        cs0 = torch.empty(ci, dtype=torch.float32, device=cuda_device)
        cs1 = torch.empty_like(cs0)
        ops.conv2d_dgrad(dy, wd, dx0, k, k, relu_mask=act, colsum=cs0)
        ops.conv2d_dgrad(dy, wd, dx1, k, k, relu_mask_bits=abits, colsum=cs1)
        torch.cuda.synchronize()
        assert torch.equal(dx1, dx0) and torch.equal(cs1, cs0)


def test_first_layer_relu_bits(ops, cuda_device):
    n, h, w = 2, 16, 72
    img = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=cuda_device)
    wt = _conv_case((1, 4, 4, 3, 64, 3), 82)[1] * 0.05
    wk1 = ops.pack_im2col_weights(dev_f32(wt, cuda_device))
    bias = dev_f32(np.full(64, -2.0, np.float32), cuda_device)
    y0 = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=cuda_device)
    y1 = torch.empty_like(y0)
    bits = torch.full((n, h, w, 2), 0x5a5a5a5a, dtype=torch.int32, device=cuda_device)
    ops.conv2d_first_fwd(img, wk1, bias, y0, 3, 3, relu=True)
    ops.conv2d_first_fwd(img, wk1, bias, y1, 3, 3, relu=True, relu_bits=bits)
    torch.cuda.synchronize()
    assert torch.equal(y1, y0)
    assert torch.equal(bits, _pack_bits(y1))
    assert 0.05 < float((y1 > 0).float().mean()) < 0.95


@pytest.mark.parametrize("shape", [(3, 10, 38, 64, 64, 3), (2, 12, 70, 64, 128, 3), (1, 6, 46, 128, 256, 3), (2, 8, 34, 64, 64, 3),
                                   (5, 6, 10, 64, 192, 1)])
def test_fused_pool_and_mask_bits_stay_inside_their_outputs(ops, cuda_device, shape):
    """Ragged maps (widths that are not multiples of the 30- / box-wide tiles, batches that do not fill the last box)
    through the fused pool epilogue and the mask-bit stores, with guard bytes around every output."""
    n, h, w, ci, co, k = shape
    x, wt, b = _conv_case(shape, 90)
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, bd = dev_bf16(x, cuda_device), dev_f32(b, cuda_device)
    y, chk_y = _guarded((n, h, w, co), torch.bfloat16, cuda_device)
    p, chk_p = _guarded((n, h // 2, w // 2, co), torch.bfloat16, cuda_device)
    i, chk_i = _guarded((n, h // 2, w // 2, co), torch.uint8, cuda_device)
    for pool_only in (False, True):
        ops.conv2d_fwd_pool(xd, wk, bd, y, p, i, k, k, relu=True, pool_only=pool_only)
        chk_y("fused-pool conv output"); chk_p("pooled tensor"); chk_i("pool indices")
    y0 = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    p0, i0 = torch.empty_like(p), torch.empty_like(i)
    ops.conv2d_fwd_pool(xd, wk, bd, y, p, i, k, k, relu=True, pool_only=False)
    ops.maxpool_fwd(y, p0, i0)
    torch.cuda.synchronize()
    assert torch.equal(p, p0) and torch.equal(i, i0)
    bits, chk_b = _guarded((n, h, w, co // 32), torch.int32, cuda_device)
    ops.conv2d_fwd(xd, wk, bd, y0, k, k, relu=True, relu_bits=bits)
    chk_b("relu bits")
    assert torch.equal(bits, _pack_bits(y0))
    dx, chk_dx = _guarded((n, h, w, ci), torch.bfloat16, cuda_device)
    dy = dev_bf16(bf16_grid(np.random.default_rng(91).standard_normal((n, h, w, co))), cuda_device)
    ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask_bits=_pack_bits(xd))
    chk_dx("dgrad with mask bits")


PAIR_SHAPES = [
    # N, H, W, Cin, Cout, k            256-column tiles run as CTA pairs (cta_group::2, M = 256) when there are >= 16 pixel tiles
    (4, 20, 72, 256, 512, 3),          # 45 pixel tiles (odd: the last pair has a partner outside the batch) x 2 channel tiles
    (8, 10, 36, 512, 512, 3),          # border tiles whose tap sets differ inside a pair (union walk)
    (2, 40, 144, 128, 256, 3),         # one channel tile
    (32, 5, 18, 128, 1024, 7),         # pixel-tile-fastest order (weight-heavy), 7x7 taps mostly padding
    (5, 24, 40, 256, 256, 1),          # 1x1
    (32, 10, 36, 512, 512, 3),         # conv5_x at B=32: 90 pixel tiles, 45 pairs per channel tile
    (32, 20, 72, 512, 256, 3),         # one 256-column channel tile, 360 pixel tiles
    (2, 64, 96, 128, 128, 3),          # haloed-slab kernel, 128-column tiles as pairs (slab_pair_kernel<128>, tuning pair = 2): 128 tiles
    (1, 60, 90, 64, 128, 3),           # slab pairs, 45 tiles (odd: one partner outside the batch); dgrad is the kx-fused slab3
]


@pytest.mark.parametrize("shape", PAIR_SHAPES)
def test_cta_pair_igemm_matches_single_cta_bit_for_bit(ops, cuda_device, shape):
    """igemm_pair_kernel: two CTAs of a cluster run two pixel tiles as one M = 256 tcgen05.mma.cta_group::2, each staging
    half of the weight tile.  Same k-step order per output element as the single-CTA kernel (extra all-padding taps of the
    union walk add zeros), so forward (plain / fused pool / mask bits), dgrad (mask bits, residual, column sums) must agree
    exactly; parity with the oracle is covered by the ordinary conv tests, which take this path by default."""
    n, h, w, ci, co, k = shape
    rng = np.random.default_rng(11)
    x = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    wt = dev_f32(bf16_grid(rng.standard_normal((k, k, ci, co)) / np.sqrt(k * k * ci)), cuda_device)
    b = dev_f32(rng.standard_normal(co) * 0.1, cuda_device)
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, co))), cuda_device)
    res = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    wk, wd = ops.pack_conv_weights(wt)
    out = {}
    try:
        for mode in (0, 2):          # 2: igemm pairs (the default, 1) plus the slab kernel's pair form
            ops.ctx.set_tuning("pair", mode)
            l0 = ops.ctx.launches
            y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
            bits = torch.zeros((n, h, w, co // 32), dtype=torch.int32, device=cuda_device)
            ops.conv2d_fwd(x, wk, b, y, k, k, relu=True, relu_bits=bits)
            y2 = torch.zeros_like(y)
            pooled = torch.zeros((n, h // 2, w // 2, co), dtype=torch.bfloat16, device=cuda_device)
            idx = torch.zeros((n, h // 2, w // 2, co), dtype=torch.uint8, device=cuda_device)
            if h % 2 == 0 and w % 2 == 0:
                ops.conv2d_fwd_pool(x, wk, b, y2, pooled, idx, k, k, relu=True)
            xbits = torch.empty((n, h, w, ci // 32), dtype=torch.int32, device=cuda_device)
            ops.relu_bits(x, xbits)
            dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
            cs = torch.empty(ci, dtype=torch.float32, device=cuda_device)
            ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask_bits=xbits, residual=res, colsum=cs)
            dx2 = torch.empty_like(dx)
            ops.conv2d_dgrad(dy, wd, dx2, k, k, relu_mask=x)
            dw = torch.empty((k, k, ci, co), dtype=torch.float32, device=cuda_device)
            ops.conv2d_wgrad(x, dy, dw, k, k)            # (wgrad_pair_kernel for the 256-column layers whose boxes all contribute)
            torch.cuda.synchronize()
            out[mode] = (y, bits, y2, pooled, idx, dx, cs, dx2, dw)
    finally:
        ops.ctx.set_tuning("pair", 1)
    names = ("fwd", "relu bits", "fwd (pool call)", "pooled", "pool idx", "dgrad", "dgrad column sums", "dgrad (bf16 mask)", "wgrad")
    for a, c, name in zip(out[0], out[2], names):
        if name == "dgrad column sums":      # per-CTA partial rows: the grid (hence the fixed summation order) differs between the modes
            np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), rtol=1e-4, atol=1e-3)
        else:
            assert torch.equal(a, c), f"{name} {shape}: pair vs single-CTA"
    # and against the oracle once more, explicitly in pair mode
    ref = T.relu(T.bias_add(T.conv2d_same(x.float().cpu(), wt.cpu()), b.cpu())).numpy()
    assert_close(host(out[2][0]), ref, TOL_BF16, f"pair conv fwd {shape}")


@pytest.mark.parametrize("shape", [(8, 20, 36, 256, 256), (4, 40, 72, 256, 128)])
def test_cta_pair_strided_conv_matches_single_cta(ops, cuda_device, shape):
    """The stride-2 implicit GEMM (input gradient of the 4x4 / stride-2 transposed conv; forward of LidCamNet's 4x4 stride-2
    conv) as CTA pairs: per-tap decimated views of dy as the A operand, same bits as single-CTA tiles, and the oracle."""
    n, h, w, ci, co = shape          # transposed conv ci -> co, x [n,h,w,ci], dy [n,2h,2w,co]
    rng = np.random.default_rng(12)
    wt = bf16_grid(rng.standard_normal((4, 4, co, ci)) / np.sqrt(4 * co))
    dy = bf16_grid(rng.standard_normal((n, 2 * h, 2 * w, co)))
    _, wd = ops.pack_deconv_weights(dev_f32(wt, cuda_device), 2)
    dyd = dev_bf16(dy, cuda_device)
    xd = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    out = {}
    try:
        for mode in (0, 1):
            ops.ctx.set_tuning("pair", mode)
            dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
            cs = torch.empty(ci, dtype=torch.float32, device=cuda_device)
            ops.deconv2d_dgrad(dyd, wd, dx, 4, 2, colsum=cs)
            dw = torch.empty((4, 4, co, ci), dtype=torch.float32, device=cuda_device)
            ops.deconv2d_wgrad(xd, dyd, dw, 4, 2)          # wgrad_pair_kernel when Cin % 256 == 0: decimated dy views as A, x boxes as B
            torch.cuda.synchronize()
            out[mode] = (dx, cs, dw)
    finally:
        ops.ctx.set_tuning("pair", 1)
    assert torch.equal(out[0][0], out[1][0])
    assert torch.equal(out[0][2], out[1][2]), "deconv wgrad: pair vs single-CTA"
    xt = torch.tensor(host(xd), requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    T.conv2d_transpose_same(xt, wtt, (2 * h, 2 * w), 2).backward(torch.tensor(dy))
    assert_close(host(out[1][2]), wtt.grad.numpy(), TOL_F32, f"pair deconv wgrad {shape}")
    np.testing.assert_allclose(out[0][1].cpu().numpy(), out[1][1].cpu().numpy(), rtol=1e-4, atol=1e-3)
    # oracle: gradient of conv2d_transpose wrt its input = stride-2 conv of dy with W[k,k,Cout,Cin] read as HWIO [k,k,co,ci]
    ref = T.conv2d_same(torch.tensor(dy), torch.tensor(wt), stride=2).numpy()
    assert_close(host(out[1][0]), ref, TOL_BF16, f"pair strided conv {shape}")
