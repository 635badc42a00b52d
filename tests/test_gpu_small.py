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


def test_conv8_skinny_fwd_dgrad_wgrad(ops, cuda_device):
    n, h, w, ci, co = 2, 5, 18, 4096, 2
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


@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, s, f32 output/grad
    (2, 5, 18, 2, 512, 4, 2, False),      # conv_t1
    (1, 6, 8, 256, 2, 16, 8, True),       # conv_t3 (logits fp32)
    (2, 4, 4, 16, 24, 4, 2, False),
])
def test_deconv_small_fwd_dgrad_wgrad(ops, cuda_device, case):
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
