"""GPU parity of the remaining op families of Network/utils/utils.py against oracle/tf_ops.py:
Atrous_Conv2D_Layer (:210-231) forward / dgrad / wgrad on the tcgen05 implicit GEMM, Resize_Bilinear with
align_corners=True (:329-330) forward / backward, Global_Avg_Pool (:312-313), Global_Max_Pool (:315-316), Avg_Pooling
(:309), Zero_Padding (:325-327), and the 4x4 stride-2 Conv2D of LidCamNet.py:28-33."""
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


@pytest.mark.parametrize("case", [(2, 20, 24, 64, 128, 3, 2), (1, 12, 36, 128, 64, 3, 4), (2, 17, 19, 64, 64, 3, 6),
                                  (1, 10, 36, 256, 256, 3, 12)])
def test_atrous_conv_fwd_dgrad_wgrad(ops, cuda_device, case):
    n, h, w, ci, co, k, rate = case
    rng = np.random.default_rng(50)
    x = bf16_grid(rng.standard_normal((n, h, w, ci)))
    wt = bf16_grid(rng.standard_normal((k, k, ci, co)) / np.sqrt(k * k * ci))
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    act = bf16_grid(rng.standard_normal((n, h, w, ci)))
    xt, wtt = torch.tensor(x, requires_grad=True), torch.tensor(wt, requires_grad=True)
    z = T.atrous_conv2d_same(xt, wtt, rate)
    y_ref = T.relu(T.bias_add(z, torch.tensor(b))).detach().numpy()
    z.backward(torch.tensor(dy))
    wk, wd = ops.pack_conv_weights(dev_f32(wt, cuda_device))
    xd, dyd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device)
    y = torch.full((n, h, w, co), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.atrous_conv2d_fwd(xd, wk, dev_f32(b, cuda_device), y, k, rate, relu=True)
    dx = torch.full((n, h, w, ci), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.atrous_conv2d_dgrad(dyd, wd, dx, k, rate, relu_mask=dev_bf16(act, cuda_device))
    dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.atrous_conv2d_wgrad(xd, dyd, dw, k, rate)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref, 1e-2, f"atrous fwd {case}")
    assert_close(host(dx), xt.grad.numpy() * (act > 0), 1e-2, f"atrous dgrad {case}")
    assert_close(host(dw), wtt.grad.numpy(), 2e-3, f"atrous wgrad {case}")


@pytest.mark.parametrize("case", [(2, 5, 18, 64, 160, 576), (1, 20, 72, 32, 160, 576), (2, 7, 9, 8, 7, 9), (1, 12, 10, 16, 5, 1),
                                  (1, 10, 36, 64, 1, 1)])
def test_resize_bilinear_align_corners_fwd_bwd(ops, cuda_device, case):
    n, h, w, c, oh, ow = case
    rng = np.random.default_rng(51)
    x = bf16_grid(rng.standard_normal((n, h, w, c)))
    dy = bf16_grid(rng.standard_normal((n, oh, ow, c)))
    xt = torch.tensor(x, requires_grad=True)
    y_ref = T.resize_bilinear_align_corners(xt, (oh, ow))
    y_ref.backward(torch.tensor(dy))
    y = torch.empty((n, oh, ow, c), dtype=torch.bfloat16, device=cuda_device)
    ops.resize_bilinear_fwd(dev_bf16(x, cuda_device), y)
    dx = torch.full((n, h, w, c), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.resize_bilinear_bwd(dev_bf16(dy, cuda_device), dx)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), 1e-2, f"resize_bilinear fwd {case}")
    assert_close(host(dx), xt.grad.numpy(), 1e-2, f"resize_bilinear bwd {case}")
    # deterministic (gather form, no atomics)
    dx2 = torch.empty_like(dx)
    ops.resize_bilinear_bwd(dev_bf16(dy, cuda_device), dx2)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2)


def test_global_and_2x2_average_pools(ops, cuda_device):
    rng = np.random.default_rng(52)
    n, h, w, c = 3, 10, 36, 520
    x = bf16_grid(rng.standard_normal((n, h, w, c)))
    xd = dev_bf16(x, cuda_device)
    y = torch.empty((n, c), dtype=torch.bfloat16, device=cuda_device)
    ops.global_avgpool_fwd(xd, y)
    dyg = bf16_grid(rng.standard_normal((n, c)))
    dx = torch.empty_like(xd)
    ops.global_avgpool_bwd(dev_bf16(dyg, cuda_device), dx)
    torch.cuda.synchronize()
    assert_close(host(y), T.global_avg_pool(torch.tensor(x)).numpy(), 1e-2, "global avg pool")
    assert_close(host(dx), np.broadcast_to(dyg[:, None, None, :] / (h * w), x.shape), 1e-2, "global avg pool grad")
    # Avg_Pooling 2x2 on a channel prefix of a wider buffer (the strided view the dense blocks use)
    ld, cu = 64, 40
    xb = bf16_grid(rng.standard_normal((2, 8, 12, ld)))
    yb = torch.zeros((2, 4, 6, ld), dtype=torch.bfloat16, device=cuda_device)
    ops.avgpool_fwd(dev_bf16(xb, cuda_device), cu, yb)
    torch.cuda.synchronize()
    ref = bf16_grid(T.avg_pool_2x2(torch.tensor(xb[..., :cu])).numpy())
    assert np.array_equal(host(yb)[..., :cu], ref) and float(yb[..., cu:].abs().max()) == 0.0
    dyb = bf16_grid(rng.standard_normal((2, 4, 6, ld)))
    dxb = torch.zeros((2, 8, 12, ld), dtype=torch.bfloat16, device=cuda_device)
    ops.avgpool_bwd(dev_bf16(dyb, cuda_device), cu, dxb)
    torch.cuda.synchronize()
    up = np.repeat(np.repeat(dyb[..., :cu], 2, axis=1), 2, axis=2) * 0.25
    assert np.array_equal(host(dxb)[..., :cu], bf16_grid(up)) and float(dxb[..., cu:].abs().max()) == 0.0


def test_global_max_pool_and_zero_padding(ops, cuda_device):
    """Global_Max_Pool (utils.py:315-316) with TF's tie-sharing gradient, Zero_Padding (utils.py:325-327) and its crop."""
    rng = np.random.default_rng(60)
    n, h, w, c = 3, 9, 14, 136
    x = bf16_grid(np.round(rng.standard_normal((n, h, w, c)) * 2) / 2)          # coarse grid: ties at the maximum do occur
    xt = torch.tensor(x, requires_grad=True)
    y_ref = T.global_max_pool(xt)
    dy = bf16_grid(rng.standard_normal((n, c)))
    y_ref.backward(torch.tensor(dy))
    xd = dev_bf16(x, cuda_device)
    y = torch.empty((n, c), dtype=torch.bfloat16, device=cuda_device)
    cnt = torch.empty((n, c), dtype=torch.int32, device=cuda_device)
    dx = torch.empty_like(xd)
    ops.global_maxpool_fwd(xd, y, cnt)
    ops.global_maxpool_bwd(dev_bf16(dy, cuda_device), xd, y, cnt, dx)
    torch.cuda.synchronize()
    assert np.array_equal(host(y), y_ref.detach().numpy())
    ties = (x == x.max(axis=(1, 2), keepdims=True)).sum(axis=(1, 2))
    assert np.array_equal(cnt.cpu().numpy(), ties) and ties.max() > 1
    assert_close(host(dx), xt.grad.numpy(), 1e-2, "global max pool grad (dy / ties on every maximal element)")
    pad = 3
    yp = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=torch.bfloat16, device=cuda_device)
    ops.zero_pad(xd, yp, pad)
    back = torch.empty_like(xd)
    ops.zero_pad(yp, back, pad, crop=True)
    torch.cuda.synchronize()
    assert np.array_equal(host(yp), T.zero_padding(torch.tensor(x), pad).numpy())
    assert torch.equal(back, xd)


@pytest.mark.parametrize("shape", [(2, 12, 20, 64, 128), (3, 10, 36, 128, 64)])
def test_conv2d_4x4_stride2_fwd_dgrad_wgrad(ops, cuda_device, shape):
    """Conv2D 4x4 / stride 2 / SAME (LidCamNet.py:28-33) on the transposed conv's kernels: forward = the strided implicit
    GEMM with bias + ReLU, input gradient = conv2d_transpose, weight gradient with x / dy swapped; vs tf.nn.conv2d."""
    n, h, w, ci, co = shape                       # (h, w) = OUTPUT size; the input is 2h x 2w
    rng = np.random.default_rng(61)
    x = bf16_grid(rng.standard_normal((n, 2 * h, 2 * w, ci)))
    wt = bf16_grid(rng.standard_normal((4, 4, ci, co)) / np.sqrt(16 * ci))
    b = rng.standard_normal(co).astype(np.float32) * 0.1
    dy = bf16_grid(rng.standard_normal((n, h, w, co)))
    xt, wtt = torch.tensor(x, requires_grad=True), torch.tensor(wt, requires_grad=True)
    z = T.bias_add(T.conv2d_same(xt, wtt, stride=2), torch.tensor(b))
    y_ref = T.relu(z).detach().numpy()
    z.backward(torch.tensor(dy))
    wk, wd = ops.pack_deconv_weights(dev_f32(wt, cuda_device), 2)
    xd, dyd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device)
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    dx = torch.empty_like(xd)
    dw = torch.full((4, 4, ci, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.conv2d_s2_fwd(xd, wd, dev_f32(b, cuda_device), y, relu=True)
    ops.conv2d_s2_dgrad(dyd, wk, dx)
    ops.conv2d_s2_wgrad(xd, dyd, dw)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref, 1e-2, f"4x4 s2 conv fwd {shape}")
    assert_close(host(dx), xt.grad.numpy(), 1e-2, f"4x4 s2 conv dgrad {shape}")
    assert_close(host(dw), wtt.grad.numpy(), 2e-3, f"4x4 s2 conv wgrad {shape}")


# ---- round 2, second batch: windowed pools, depthwise conv, sigmoid / swish, squeeze-excite multiply (csrc/opfam.cu) -----

@pytest.mark.parametrize("case", [(2, 20, 72, 64, 20, 72, 20, 72), (2, 20, 72, 64, 10, 36, 10, 36), (1, 20, 72, 128, 7, 24, 7, 24),
                                  (2, 20, 72, 64, 4, 12, 4, 12), (3, 9, 11, 40, 3, 3, 2, 2), (2, 8, 12, 24, 2, 2, 2, 2)])
def test_avg_pooling_windows_fwd_bwd(ops, cuda_device, case):
    """Avg_Pooling(kh, kw, stride_h, stride_w, VALID) (utils.py:309): PSPNet's pyramid windows at 20x72 (PSPNet.py:546-567),
    an overlapping 3x3 / stride-2 window, and the 2x2 / stride-2 one."""
    n, h, w, c, kh, kw, sh, sw = case
    rng = np.random.default_rng(61)
    x = bf16_grid(rng.standard_normal((n, h, w, c)))
    ref = T.avg_pool_valid(torch.tensor(x), kh, kw, sh, sw)
    y = torch.empty(tuple(ref.shape), dtype=torch.bfloat16, device=cuda_device)
    ops.avgpool_window_fwd(dev_bf16(x, cuda_device), y, kh, kw, sh, sw)
    dy = bf16_grid(rng.standard_normal(tuple(ref.shape)))
    xr = torch.tensor(x, requires_grad=True)
    T.avg_pool_valid(xr, kh, kw, sh, sw).backward(torch.tensor(dy))
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=cuda_device)
    ops.avgpool_window_bwd(dev_bf16(dy, cuda_device), dx, kh, kw, sh, sw)
    torch.cuda.synchronize()
    assert_close(host(y), ref.numpy(), 1e-2, f"avg pool {case}")
    assert_close(host(dx), xr.grad.numpy(), 1e-2, f"avg pool grad {case}")


@pytest.mark.parametrize("case", [(2, 82, 290, 64, 3, 2, "VALID"), (2, 40, 144, 64, 3, 2, "SAME"), (1, 7, 9, 8, 3, 2, "SAME"),
                                  (2, 9, 11, 16, 3, 1, "SAME"), (2, 8, 12, 32, 2, 2, "VALID")])
def test_max_pooling_windows_fwd_bwd_bit_exact(ops, cuda_device, case):
    """Max_Pooling(3, 3, stride 2) VALID after Zero_Padding (PSPNet.py:32-34) and SAME (PSPNet.py:190): values, first-max
    indices and the gradient routing (overlapping windows accumulate) are exact; inputs are ReLU outputs, so ties (zeros)
    are common and the first-max rule is what is tested."""
    n, h, w, c, k, s, pad = case
    rng = np.random.default_rng(62)
    x = bf16_grid(np.maximum(rng.standard_normal((n, h, w, c)), 0))
    y_ref, idx_ref = T.max_pool_general(torch.tensor(x), k, k, s, pad)
    y = torch.empty(tuple(y_ref.shape), dtype=torch.bfloat16, device=cuda_device)
    idx = torch.empty(tuple(y_ref.shape), dtype=torch.uint8, device=cuda_device)
    ops.maxpool_general_fwd(dev_bf16(x, cuda_device), y, idx, k, k, s, same=(pad == "SAME"))
    dy = bf16_grid(rng.integers(-4, 5, tuple(y_ref.shape)).astype(np.float32))      # small integers: sums of up to 4 are exact in bf16
    dx_ref = T.max_pool_general_grad(torch.tensor(dy), idx_ref, (h, w), k, k, s, pad)
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=cuda_device)
    ops.maxpool_general_bwd(dev_bf16(dy, cuda_device), idx, dx, k, k, s, same=(pad == "SAME"))
    torch.cuda.synchronize()
    assert np.array_equal(host(y), y_ref.numpy())
    assert np.array_equal(idx.cpu().numpy(), idx_ref.numpy())
    assert np.array_equal(host(dx), dx_ref.numpy())


@pytest.mark.parametrize("case", [(2, 20, 36, 64, 3, 1, 1), (2, 21, 37, 32, 3, 2, 1), (1, 20, 36, 728, 3, 1, 2), (2, 12, 18, 96, 5, 2, 1),
                                  (1, 16, 20, 2304, 3, 1, 4)])
def test_depthwise_conv_fwd_dgrad_wgrad(ops, cuda_device, case):
    """tf.nn.depthwise_conv2d(x, filter, stride, SAME, rate) (DeepLabv3Plus.py:49 SepConv_BN: 3x3, stride 1 / 2, atrous rates;
    EfficientNet.py:173 MBConv: 3x3 / 5x5, stride 1 / 2): forward with bias, input gradient, filter gradient."""
    n, h, w, c, k, s, r = case
    rng = np.random.default_rng(63)
    x = bf16_grid(rng.standard_normal((n, h, w, c)))
    wt = bf16_grid(rng.standard_normal((k, k, c)) / k)
    b = rng.standard_normal(c).astype(np.float32) * 0.1
    xr, wr = torch.tensor(x, requires_grad=True), torch.tensor(wt, requires_grad=True)
    ref = T.depthwise_conv2d_same(xr, wr, s, r) + torch.tensor(b)
    dy = bf16_grid(rng.standard_normal(tuple(ref.shape)))
    ref.backward(torch.tensor(dy))
    xd, wd_, dyd = dev_bf16(x, cuda_device), dev_f32(wt, cuda_device), dev_bf16(dy, cuda_device)
    y = torch.empty(tuple(ref.shape), dtype=torch.bfloat16, device=cuda_device)
    ops.depthwise_conv2d_fwd(xd, wd_, dev_f32(b, cuda_device), y, stride=s, rate=r)
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=cuda_device)
    ops.depthwise_conv2d_dgrad(dyd, wd_, dx, stride=s, rate=r)
    dw = torch.empty((k, k, c), dtype=torch.float32, device=cuda_device)
    ops.depthwise_conv2d_wgrad(xd, dyd, dw, stride=s, rate=r)
    dw2 = dw.clone()
    ops.depthwise_conv2d_wgrad(xd, dyd, dw2, stride=s, rate=r, accumulate=True)
    torch.cuda.synchronize()
    assert_close(host(y), ref.detach().numpy(), 1e-2, f"depthwise fwd {case}")
    assert_close(host(dx), xr.grad.numpy(), 1e-2, f"depthwise dgrad {case}")
    assert_close(host(dw), wr.grad.numpy(), 2e-3, f"depthwise wgrad {case}")
    assert_close(host(dw2), 2 * wr.grad.numpy(), 2e-3, f"depthwise wgrad accumulate {case}")


def test_sigmoid_swish_and_squeeze_excite_multiply(ops, cuda_device):
    """tf.sigmoid / swish (EfficientNet.py's activation) forward + backward and the SE multiply x * s[n, c] with ds = sum_hw dy x."""
    rng = np.random.default_rng(64)
    n, h, w, c = 3, 10, 18, 96
    x = bf16_grid(rng.standard_normal((n, h, w, c)) * 2)
    dy = bf16_grid(rng.standard_normal((n, h, w, c)))
    xd, dyd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device)
    for kind, fn in (("sigmoid", torch.sigmoid), ("swish", T.swish)):
        xr = torch.tensor(x, requires_grad=True)
        ref = fn(xr)
        ref.backward(torch.tensor(dy))
        y, dx = torch.empty_like(xd), torch.empty_like(xd)
        ops.activation_fwd(xd, y, kind)
        ops.activation_bwd(xd, dyd, dx, kind)
        torch.cuda.synchronize()
        assert_close(host(y), ref.detach().numpy(), 1e-2, kind)
        assert_close(host(dx), xr.grad.numpy(), 1e-2, kind + " grad")
    s = bf16_grid(1 / (1 + np.exp(-rng.standard_normal((n, c)))))
    xr, sr = torch.tensor(x, requires_grad=True), torch.tensor(s, requires_grad=True)
    ref = T.channel_scale(xr, sr)
    ref.backward(torch.tensor(dy))
    sd = dev_bf16(s, cuda_device)
    y, dx = torch.empty_like(xd), torch.empty_like(xd)
    ds = torch.empty((n, c), dtype=torch.float32, device=cuda_device)
    ops.channel_scale_fwd(xd, sd, y)
    ops.channel_scale_bwd(xd, sd, dyd, dx, ds)
    torch.cuda.synchronize()
    assert_close(host(y), ref.detach().numpy(), 1e-2, "channel scale")
    assert_close(host(dx), xr.grad.numpy(), 1e-2, "channel scale dx")
    assert_close(host(ds), sr.grad.numpy(), 2e-3, "channel scale ds")


def test_op_family_argument_errors(ops, cuda_device):
    """Unsupported shapes are errors with a message, never a silent fallback."""
    from semanticsegmentation_tensorflow_b200._lib import SegkError
    x = torch.zeros((1, 8, 8, 16), dtype=torch.bfloat16, device=cuda_device)
    y = torch.zeros_like(x)
    w = torch.zeros((3, 3, 16), dtype=torch.float32, device=cuda_device)
    with pytest.raises(SegkError, match="rate > 1 needs stride 1"):
        ops.depthwise_conv2d_fwd(x, w, None, y[:, :4, :4], stride=2, rate=2)
    with pytest.raises(SegkError, match="C %% 8|C % 8"):
        ops.avgpool_window_fwd(x[..., :12].contiguous(), y[..., :12].contiguous(), 2, 2, 2, 2)
    with pytest.raises(SegkError, match="larger than"):
        ops.avgpool_window_fwd(x, y, 9, 9, 1, 1)
    with pytest.raises(SegkError, match="byte"):
        ops.maxpool_general_fwd(torch.zeros((1, 20, 20, 8), dtype=torch.bfloat16, device=cuda_device),
                                torch.zeros((1, 1, 1, 8), dtype=torch.bfloat16, device=cuda_device),
                                torch.zeros((1, 1, 1, 8), dtype=torch.uint8, device=cuda_device), 17, 17, 1)
    with pytest.raises(KeyError):
        ops.activation_fwd(x, y, "gelu")
    torch.cuda.synchronize()
