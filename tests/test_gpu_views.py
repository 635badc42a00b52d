"""GPU parity: channel-slice views (segk_set_pitch) = zero-copy Concat (utils.py:332).

A view is a channel slice [c0, c0+C) of a wider NHWC buffer.  Every call that accepts one must give the bits of the same
call on a dense copy (the arithmetic is the same; only the TMA descriptor / row stride differs), must leave the other
channels of the wide buffer untouched, and any call that does not accept views must fail while a pitch is pending."""
import numpy as np
import pytest
import torch

from tests.gpu_util import bf16_grid, dev_bf16, dev_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


SENT = 123.0     # sentinel in the channels a view does not own


def _wide(shape, c0, ctot, dev, src=None):
    """(wide buffer filled with the sentinel, its channel slice [c0, c0 + C) -- holding src when given)."""
    n, h, w, c = shape
    wide = torch.full((n, h, w, ctot), SENT, dtype=torch.bfloat16, device=dev)
    view = wide[..., c0:c0 + c]
    if src is not None:
        view.copy_(src)
    return wide, view


def _others_untouched(wide, c0, c):
    keep = torch.ones(wide.shape[3], dtype=torch.bool, device=wide.device)
    keep[c0:c0 + c] = False
    return bool((wide[..., keep].float() == SENT).all())


VIEW_SHAPES = [
    # N, H, W, Cin, Cout, k, c0, Ctot          (which kernel the forward / dgrad / wgrad of the shape take)
    (2, 16, 24, 64, 64, 3, 64, 128),          # igemm, wgrad_kernel
    (2, 64, 96, 64, 64, 3, 64, 128),          # slab3 fwd / dgrad, wslab<64,1>
    (2, 64, 96, 128, 128, 3, 128, 256),       # slab<128>, wslab<128,2>
    (2, 20, 72, 256, 256, 3, 0, 512),         # igemm<256>, first slice
    (3, 10, 12, 128, 64, 1, 192, 256),        # 1x1
]


@pytest.mark.parametrize("shape", VIEW_SHAPES)
def test_conv_fwd_pool_into_a_view_and_gradients_from_a_view(ops, cuda_device, shape):
    n, h, w, ci, co, k, c0, ctot = shape
    rng = np.random.default_rng(1)
    x = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    wt = dev_f32(bf16_grid(rng.standard_normal((k, k, ci, co)) / np.sqrt(k * k * ci)), cuda_device)
    b = dev_f32(rng.standard_normal(co) * 0.1, cuda_device)
    wk, wd = ops.pack_conv_weights(wt)
    # forward: plain and with the pool in the epilogue, dense vs into the slice
    y0 = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(x, wk, b, y0, k, k, relu=True)
    wide, yv = _wide((n, h, w, co), c0, ctot, cuda_device)
    ops.conv2d_fwd(x, wk, b, yv, k, k, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(yv, y0), f"conv fwd into a view {shape}"
    assert _others_untouched(wide, c0, co)
    p0 = torch.empty((n, h // 2, w // 2, co), dtype=torch.bfloat16, device=cuda_device)
    i0 = torch.empty((n, h // 2, w // 2, co), dtype=torch.uint8, device=cuda_device)
    ops.conv2d_fwd_pool(x, wk, b, y0, p0, i0, k, k, relu=True)
    wide, yv = _wide((n, h, w, co), c0, ctot, cuda_device)
    p1, i1 = torch.empty_like(p0), torch.empty_like(i0)
    ops.conv2d_fwd_pool(x, wk, b, yv, p1, i1, k, k, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(yv, y0) and torch.equal(p1, p0) and torch.equal(i1, i0), f"conv + pool into a view {shape}"
    assert _others_untouched(wide, c0, co)
    # the pool alone, reading a view
    p2, i2 = torch.empty_like(p0), torch.empty_like(i0)
    ops.maxpool_fwd(yv, p2, i2)
    torch.cuda.synchronize()
    assert torch.equal(p2, p0) and torch.equal(i2, i0)
    # gradients: dy is a slice of a wider gradient tensor
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, co))), cuda_device)
    _, dyv = _wide((n, h, w, co), c0, ctot, cuda_device, src=dy)
    dx0 = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    dx1 = torch.empty_like(dx0)
    ops.conv2d_dgrad(dy, wd, dx0, k, k, relu_mask=x)
    ops.conv2d_dgrad(dyv, wd, dx1, k, k, relu_mask=x)
    dw0 = torch.empty((k, k, ci, co), dtype=torch.float32, device=cuda_device)
    dw1 = torch.empty_like(dw0)
    ops.conv2d_wgrad(x, dy, dw0, k, k)
    ops.conv2d_wgrad(x, dyv, dw1, k, k)
    db0 = torch.empty(co, dtype=torch.float32, device=cuda_device)
    db1 = torch.empty_like(db0)
    ops.bias_grad(dy, db0)
    ops.bias_grad(dyv, db1)
    torch.cuda.synchronize()
    assert torch.equal(dx1, dx0), f"dgrad from a view {shape}"
    assert torch.equal(dw1, dw0), f"wgrad from a view {shape}"
    assert torch.equal(db1, db0), f"bias_grad from a view {shape}"


@pytest.mark.parametrize("shape", [(2, 8, 12, 128, 64, 0, 128), (1, 20, 36, 256, 256, 256, 512), (2, 5, 9, 512, 512, 512, 1024)])
def test_deconv_into_a_view_and_gradients_from_a_view(ops, cuda_device, shape):
    n, h, w, ci, co, c0, ctot = shape
    rng = np.random.default_rng(2)
    x = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, ci))), cuda_device)
    wt = dev_f32(bf16_grid(rng.standard_normal((4, 4, co, ci)) / np.sqrt(4 * ci)), cuda_device)
    wk, wd = ops.pack_deconv_weights(wt, 2)
    y0 = torch.empty((n, 2 * h, 2 * w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.deconv2d_fwd(x, wk, None, y0, 4, 2)
    wide, yv = _wide((n, 2 * h, 2 * w, co), c0, ctot, cuda_device)
    ops.deconv2d_fwd(x, wk, None, yv, 4, 2)
    torch.cuda.synchronize()
    assert torch.equal(yv, y0), f"deconv fwd into a view {shape}"
    assert _others_untouched(wide, c0, co)
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, 2 * h, 2 * w, co))), cuda_device)
    _, dyv = _wide((n, 2 * h, 2 * w, co), c0, ctot, cuda_device, src=dy)
    dx0 = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    dx1 = torch.empty_like(dx0)
    ops.deconv2d_dgrad(dy, wd, dx0, 4, 2, relu_mask=x)
    ops.deconv2d_dgrad(dyv, wd, dx1, 4, 2, relu_mask=x)
    dw0 = torch.empty((4, 4, co, ci), dtype=torch.float32, device=cuda_device)
    dw1 = torch.empty_like(dw0)
    ops.deconv2d_wgrad(x, dy, dw0, 4, 2)
    ops.deconv2d_wgrad(x, dyv, dw1, 4, 2)
    torch.cuda.synchronize()
    assert torch.equal(dx1, dx0), f"deconv dgrad from a view {shape}"
    assert torch.equal(dw1, dw0), f"deconv wgrad from a view {shape}"


def test_maxpool_bwd_into_a_view_adds_the_second_path_and_masks_the_sum(ops, cuda_device):
    """The skip tensor of a zero-copy Concat: its gradient slice already holds the Concat's share; the pool's backward adds
    its own and applies the ReluGrad of the activation slice to the sum -- the same bits as the dense call."""
    n, h, w, c, c0, ctot = 2, 16, 24, 64, 64, 128
    rng = np.random.default_rng(3)
    act = dev_bf16(bf16_grid(np.maximum(rng.standard_normal((n, h, w, c)), 0)), cuda_device)
    res = dev_bf16(bf16_grid(rng.standard_normal((n, h, w, c))), cuda_device)
    dy = dev_bf16(bf16_grid(rng.standard_normal((n, h // 2, w // 2, c))), cuda_device)
    pooled = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=cuda_device)
    idx = torch.empty((n, h // 2, w // 2, c), dtype=torch.uint8, device=cuda_device)
    ops.maxpool_fwd(act, pooled, idx)
    dx0 = res.clone()
    ops.maxpool_bwd(dy, idx, dx0, act=act, residual=dx0)
    _, actv = _wide((n, h, w, c), c0, ctot, cuda_device, src=act)
    gwide, dxv = _wide((n, h, w, c), c0, ctot, cuda_device, src=res)
    ops.maxpool_bwd(dy, idx, dxv, act=actv, residual=dxv)
    torch.cuda.synchronize()
    assert torch.equal(dxv, dx0)
    assert _others_untouched(gwide, c0, c)


def test_a_pending_pitch_fails_calls_that_take_dense_tensors_only(ops, cuda_device):
    from semanticsegmentation_tensorflow_b200._lib import SegkError
    x = torch.zeros((1, 8, 8, 64), dtype=torch.bfloat16, device=cuda_device)
    y = torch.empty_like(x)
    ops._call("segk_set_pitch", 128, 0)
    with pytest.raises(SegkError, match="pitch"):
        ops.dropout(x, y, 0.5, 1)
    ops.dropout(x, y, 0.5, 1)          # the failed call consumed the pitch: the context is usable again
    with pytest.raises(SegkError, match="multiples of 8"):
        ops._call("segk_set_pitch", 100, 0)
    with pytest.raises(ValueError, match="channel-slice"):
        ops.maxpool_fwd(x.permute(0, 2, 1, 3), y[:, :4, :4], torch.empty((1, 4, 4, 64), dtype=torch.uint8, device=cuda_device))
    torch.cuda.synchronize()
