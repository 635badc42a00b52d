"""GPU parity at the EXACT layer shapes of the benchmarked configuration (BASELINE
configs[1]: FCN-8s 160x576, fc = 4096), batch 2, each tensor compared in full against oracle/tf_ops.py:

  conv6   7x7  512 -> 4096 @ 5x18   (FCN.py:78)   K = 25 088, split-K dgrad, pixel-tile-fastest order
  conv7   1x1 4096 -> 4096 @ 5x18   (FCN.py:82)
  conv_t3 16x16 s8 256 -> 2 @ 20x72 (FCN.py:101-107)  patch-space GEMM + col2im, fp32 logits

and one whole-network pass at fc = 4096, 160x576, He init: every one of the 40 gradient tensors with a
HARD floor (cosine >= 0.99, rel err <= 0.15) that does not widen with the oracle's own bf16-vs-fp32 noise.
Inputs are on the bf16 grid, so the differences are accumulation order (fp32 both sides) and the
final store: 1e-2 of the tensor max for bf16 outputs, 2e-3 for fp32 outputs."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from oracle.fcn_oracle import FCN8sOracle, init_variables, synthetic_batch
from tests.gpu_util import assert_close, bf16_grid, cosine, dev_bf16, dev_f32, host, rel_err

pytestmark = pytest.mark.gpu

N = 2
TOL_BF16 = 1e-2
TOL_F32 = 2e-3


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


@pytest.mark.parametrize("layer", [("conv6", 7, 512, 4096), ("conv7", 1, 4096, 4096)])
def test_fc_layers_fwd_dgrad_wgrad_at_baseline_shape(ops, cuda_device, layer):
    name, k, ci, co = layer
    h, w = 5, 18
    rng = np.random.default_rng(60)
    x = bf16_grid(np.maximum(rng.standard_normal((N, h, w, ci)), 0))           # post-ReLU / post-pool input
    wt = bf16_grid(rng.standard_normal((k, k, ci, co), dtype=np.float32) * np.float32(math.sqrt(2.0 / (k * k * ci))))
    b = (rng.standard_normal(co) * 0.1).astype(np.float32)
    dy = bf16_grid(rng.standard_normal((N, h, w, co)))
    xt = torch.tensor(x, requires_grad=True)
    wtt = torch.tensor(wt, requires_grad=True)
    y_ref = T.relu(T.bias_add(T.conv2d_same(xt, wtt), torch.tensor(b)))
    # backward through the pre-activation (ReluGrad of THIS layer is applied by its consumer's dgrad)
    z = T.conv2d_same(xt, wtt)
    z.backward(torch.tensor(dy))
    wd32 = dev_f32(wt, cuda_device)
    wk, wd = ops.pack_conv_weights(wd32)
    xd, dyd = dev_bf16(x, cuda_device), dev_bf16(dy, cuda_device)
    y = torch.empty((N, h, w, co), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_fwd(xd, wk, dev_f32(b, cuda_device), y, k, k, relu=True)
    torch.cuda.synchronize()
    assert_close(host(y), y_ref.detach().numpy(), TOL_BF16, f"{name} fwd")
    # dgrad with the producer's ReLU mask and the dropout 1/keep scale (FCN.py:79,83 backward)
    act = bf16_grid(rng.standard_normal((N, h, w, ci)))
    ref_dx = xt.grad.numpy() * (act > 0) * 1.25
    dx = torch.empty((N, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_dgrad(dyd, wd, dx, k, k, relu_mask=dev_bf16(act, cuda_device), scale=1.25)
    torch.cuda.synchronize()
    assert_close(host(dx), ref_dx, TOL_BF16, f"{name} dgrad")
    dw = torch.full((k, k, ci, co), 7.0, dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(xd, dyd, dw, k, k)
    torch.cuda.synchronize()
    assert_close(host(dw), wtt.grad.numpy(), TOL_F32, f"{name} wgrad")


def test_conv_t3_at_baseline_shape(ops, cuda_device):
    n, h, w, ci, co, k, s = N, 20, 72, 256, 2, 16, 8
    rng = np.random.default_rng(61)
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
    assert y.shape == (n, 160, 576, 2)
    assert_close(host(y), y_ref.detach().numpy(), 1e-4, "conv_t3 fwd (fp32 logits)")
    dy = (rng.standard_normal((n, h * s, w * s, co)) / (n * h * s * w * s)).astype(np.float32)   # dlogits scale
    y_ref.backward(torch.tensor(dy))
    P = torch.empty((n, h, w, e), dtype=torch.bfloat16, device=cuda_device)
    ops.deconv_patch_gather(dev_f32(dy, cuda_device), P, k, s)
    dw = torch.empty((1, 1, e, ci), dtype=torch.float32, device=cuda_device)
    ops.conv2d_wgrad(P, xd, dw, 1, 1)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_dgrad(P, wd, dx, 1, 1)
    db = torch.empty(co, dtype=torch.float32, device=cuda_device)
    ops.bias_grad(dev_f32(dy, cuda_device), db)
    torch.cuda.synchronize()
    # dlogits is rounded to bf16 inside the patch tensor: 2^-9 relative per element
    assert_close(host(dw).reshape(k, k, co, ci), wtt.grad.numpy(), 5e-3, "conv_t3 wgrad")
    assert_close(host(dx), xt.grad.numpy(), TOL_BF16, "conv_t3 dgrad")
    np.testing.assert_allclose(host(db), dy.sum(axis=(0, 1, 2)), rtol=1e-4, atol=1e-9)


@pytest.fixture(scope="module")
def full_net(cuda_device):
    """FCN-8s with the real fc = 4096 head at 160x576, batch 2, He init (every layer numerically visible)."""
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    variables = init_variables(cin=3, ncls=2, fc=4096, seed=1234, init="he")
    x, lab = synthetic_batch(N, 160, 576, seed=0, road_shaped=True)
    x = (x // 32).astype(np.uint8)                                  # O(1) activations under He init
    net = FCN(torch.as_tensor(x).to(cuda_device), 1.0, 2, variables=variables)
    return net, variables, x, lab


def test_full_network_fc4096_all_gradients_hard_floor(cuda_device, full_net):
    net, variables, x, lab = full_net
    net.keep_prepool = True          # conv5_3 is compared below (default: only its pooled output is stored)
    net.forward()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    net.side.join()
    net.wside.join()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True, bf16_grads=True)
    loss_ref, logits_ref, grads_ref = orc.loss_and_grads(x, lab)
    assert abs(float(loss) - loss_ref) <= 2e-3 * abs(loss_ref), (float(loss), loss_ref)
    lg = net.logits.cpu().numpy()
    assert rel_err(lg, logits_ref.numpy()) <= 2e-2
    names = {"conv_t1": "fuse_1", "conv_t2": "fuse_2", "conv_t3": "logits"}
    for name in ("conv5_3", "pool5", "conv6", "conv7", "conv_t1", "conv_t2"):
        e = rel_err(net.act[name].float().cpu().numpy(), orc.acts[names.get(name, name)].detach().numpy())
        assert e <= 2e-2, f"activation {name}: rel err {e:.3e}"
    rows = []
    for name in net.vars.slots:
        g = net.vars.grad(name).cpu().numpy()
        r = grads_ref[name].numpy()
        rows.append((name, rel_err(g, r), cosine(g, r)))
    for name, e, c in rows:
        print(f"   {name:20s} rel {e:.3e} cos {c:.6f}")
    bad = [(n, e, c) for n, e, c in rows if not (c >= 0.99 and e <= 0.15)]
    assert not bad, f"gradients below the hard floor (cosine >= 0.99, rel err <= 0.15): {bad}"
    # the big matrices of the benchmarked configuration must be much better than the floor
    for name in ("conv6/weights", "conv7/weights", "conv_t3/weights", "conv_t2/weights", "conv5_3/weights"):
        e, c = next((e, c) for n, e, c in rows if n == name)
        assert c >= 0.999 and e <= 5e-2, (name, e, c)
    # argmax agreement on this He-init forward.  At initialisation 1 % of the pixels sit within a few per cent of a
    # tie: >= 99.9 % holds on the pixels whose logit margin exceeds 1 % of mean|logit|; the RAW figure
    # is asserted at >= 99.5 % here and at >= 99.9 % after training (test_gpu_fcn.py, where margins are real).
    lr = logits_ref.numpy()
    pred_ref = lr.argmax(-1)
    agree = net.pred_u8.cpu().numpy() == pred_ref
    sel = np.abs(lr[..., 1] - lr[..., 0]) >= 0.01 * np.abs(lr).mean()
    print(f"argmax agreement (He init, step 0): raw {agree.mean():.5f}, margin-conditioned {agree[sel].mean():.5f} on {sel.mean():.2%} of pixels")
    assert agree[sel].mean() >= 0.999 and agree.mean() >= 0.995
