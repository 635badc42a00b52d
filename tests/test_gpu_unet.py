"""GPU parity: the shared-helper kernels (BN-affine folding, concat copies, pool backward with a second
gradient path) and the U-Net-style builder (BASELINE configs[2]) vs the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from oracle.fcn_oracle import synthetic_batch
from oracle.graph_oracle import UNetOracle, unet_init
from tests.gpu_util import assert_close, bf16_grid, cosine, dev_bf16, dev_f32, host, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from semanticsegmentation_tensorflow_b200.ops import Ops
    return Ops(cuda_device)


def test_channel_copy_concat_and_split(ops, cuda_device):
    rng = np.random.default_rng(0)
    a = bf16_grid(rng.standard_normal((2, 5, 7, 64)))
    b = bf16_grid(rng.standard_normal((2, 5, 7, 128)))
    cat = torch.zeros((2, 5, 7, 192), dtype=torch.bfloat16, device=cuda_device)
    ops.channel_copy(dev_bf16(a, cuda_device), 0, cat, 0, 64)
    ops.channel_copy(dev_bf16(b, cuda_device), 0, cat, 64, 128)
    torch.cuda.synchronize()
    assert np.array_equal(host(cat), np.concatenate([a, b], axis=3))          # bit-exact copy
    # gradient split: slice + ReLU mask + accumulate into an existing gradient
    g = bf16_grid(rng.standard_normal((2, 5, 7, 192)))
    old = bf16_grid(rng.standard_normal((2, 5, 7, 128)))
    dst = dev_bf16(old, cuda_device)
    ops.channel_copy(dev_bf16(g, cuda_device), 64, dst, 0, 128, mask=dev_bf16(b, cuda_device), accumulate=True)
    torch.cuda.synchronize()
    assert np.array_equal(host(dst), bf16_grid((g[..., 64:] + old) * (b > 0)))


def test_bn_affine_helpers(ops, cuda_device):
    rng = np.random.default_rng(1)
    w = rng.standard_normal((3, 3, 64, 128)).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(128)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(128)).astype(np.float32)
    s = 1.0 / math.sqrt(1.001)
    out = ops.scale_columns(dev_f32(w, cuda_device), dev_f32(gamma, cuda_device), s)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(out), w * gamma * s, rtol=1e-6)
    rows = 5000
    dz = bf16_grid(rng.standard_normal((rows, 128)))
    y = bf16_grid(rng.standard_normal((rows, 128)))
    dg = torch.empty(128, dtype=torch.float32, device=cuda_device)
    ws = torch.empty(8 << 20, dtype=torch.uint8, device=cuda_device)
    ops.bn_gamma_grad(dev_bf16(dz, cuda_device), dev_bf16(y, cuda_device), dev_f32(beta, cuda_device),
                      dev_f32(gamma, cuda_device), dg, ws)
    torch.cuda.synchronize()
    ref = (dz.astype(np.float64) * (y - beta)).sum(0) / gamma
    np.testing.assert_allclose(host(dg), ref, rtol=1e-3, atol=1e-2)
    # d(beta) = BiasAddGrad of the same dz out of the same pass
    dg2 = torch.empty_like(dg)
    db = torch.empty(128, dtype=torch.float32, device=cuda_device)
    ops.bn_gamma_grad(dev_bf16(dz, cuda_device), dev_bf16(y, cuda_device), dev_f32(beta, cuda_device),
                      dev_f32(gamma, cuda_device), dg2, ws, dbeta=db)
    torch.cuda.synchronize()
    np.testing.assert_allclose(host(dg2), ref, rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(host(db), dz.astype(np.float64).sum(0), rtol=1e-3, atol=1e-2)


def test_maxpool_bwd_with_second_gradient_path(ops, cuda_device):
    rng = np.random.default_rng(2)
    x = bf16_grid(np.maximum(rng.standard_normal((2, 8, 12, 64)), 0))
    y_ref, idx_ref = T.max_pool_2x2_with_argmax(x)
    xd = dev_bf16(x, cuda_device)
    y = torch.empty((2, 4, 6, 64), dtype=torch.bfloat16, device=cuda_device)
    idx = torch.empty((2, 4, 6, 64), dtype=torch.uint8, device=cuda_device)
    ops.maxpool_fwd(xd, y, idx)
    dy = bf16_grid(rng.standard_normal(y_ref.shape))
    other = bf16_grid(rng.standard_normal(x.shape) * (x > 0))
    dx = dev_bf16(other, cuda_device)                                          # in place: residual aliases dx
    ops.maxpool_bwd(dev_bf16(dy, cuda_device), idx, dx, act=xd, residual=dx)
    torch.cuda.synchronize()
    ref = bf16_grid((T.max_pool_2x2_grad(dy, idx_ref, (8, 12)) + other) * (x > 0))
    assert np.array_equal(host(dx), ref)


N, H, W = 2, 128, 192


def _build(cuda_device, init):
    from semanticsegmentation_tensorflow_b200.graph import UNet
    variables = unet_init(3, 2, seed=1234, init=init)
    rng = np.random.default_rng(5)
    for k in variables:                      # exercise non-trivial BN affine parameters
        if k.endswith("gamma"):
            variables[k] = (1 + 0.1 * rng.standard_normal(variables[k].shape)).astype(np.float32)
        if k.endswith("beta"):
            variables[k] = (0.05 * rng.standard_normal(variables[k].shape)).astype(np.float32)
    x, lab = synthetic_batch(N, H, W, seed=0, road_shaped=True)
    if init == "he":
        x = (x // 32).astype(np.uint8)
    net = UNet(torch.as_tensor(x).to(cuda_device), 2, variables=variables)
    net.keep_prepool = True       # every activation is compared (default: a conv read only by its pool stores the pooled tensor only)
    return net, variables, x, lab


def test_unet_graph_matches_oracle_definition():
    from semanticsegmentation_tensorflow_b200.graph import graph_variable_shapes, unet_nodes
    from oracle.graph_oracle import unet_variable_shapes
    assert list(graph_variable_shapes(unet_nodes(2), 3).items()) == list(unet_variable_shapes(3, 2).items())


@pytest.mark.parametrize("init", ["he", "ref"])
def test_unet_forward_and_gradients(cuda_device, init):
    net, variables, x, lab = _build(cuda_device, init)
    pred, logits = net.create()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = UNetOracle(variables, bf16_storage=True, bf16_grads=True)
    loss_ref, logits_ref, grads_ref = orc.loss_and_grads(x, lab)
    _, _, grads_f32 = UNetOracle(variables, bf16_storage=False).loss_and_grads(x, lab)
    for name, t in net.act.items():
        e = rel_err(t.float().cpu().numpy(), orc.acts[name].detach().numpy())
        assert e <= 2e-2, f"[{init}] activation {name}: rel err {e:.3e}"
    assert abs(float(loss) - loss_ref) <= 2e-3 * abs(loss_ref)
    worst = ("", 0.0, 1.0)
    for name in net.vars.slots:
        g, r, f = net.vars.grad(name).cpu().numpy(), grads_ref[name].numpy(), grads_f32[name].numpy()
        e, c = rel_err(g, r), cosine(g, r)
        tol_e, tol_c = max(5e-2, 3 * rel_err(r, f)), min(0.999, 1 - 9 * (1 - cosine(r, f)))
        tol_e, tol_c = min(tol_e, 0.15), max(tol_c, 0.99)            # hard floor, whatever the oracle's own bf16 noise
        assert e <= tol_e and c >= tol_c, f"[{init}] grad {name}: rel {e:.3e} (tol {tol_e:.3e}) cos {c:.6f} (tol {tol_c:.6f})"
        if e > worst[1]:
            worst = (name, e, c)
    print(f"[{init}] unet worst grad {worst[0]}: rel {worst[1]:.3e} cos {worst[2]:.6f}")


def test_unet_zero_copy_concat_matches_the_copy_form_bit_for_bit(cuda_device):
    """Concat (utils.py:332) without copies: the transposed convs and the encoder skip convs write into channel slices of
    the concat buffers and their gradients are read in place (GraphNet._concat_slots).  Same kernels, same arithmetic:
    logits and every gradient must equal the segk_channel_copy form exactly."""
    from semanticsegmentation_tensorflow_b200.graph import UNet
    net, variables, x, lab = _build(cuda_device, "he")
    assert sorted(net.slot) == sorted(["unpool%d" % i for i in range(1, 6)] + ["conv13", "conv10", "conv7", "conv4", "conv2"])
    for t, (cat, off) in net.slot.items():
        assert net.act[t].data_ptr() == net.act[cat].data_ptr() + 2 * off and not net.act[t].is_contiguous()
    ref = UNet(torch.as_tensor(x).to(cuda_device), 2, variables=variables, zero_copy_concat=False)
    ref.keep_prepool = True
    assert not ref.slot
    labd = torch.as_tensor(lab).to(cuda_device)
    for m in (net, ref):
        m.create()
        m.loss(labd, with_grad=True)
        m.backward()
    torch.cuda.synchronize()
    assert torch.equal(net.logits, ref.logits)
    for name in net.act:
        assert torch.equal(net.act[name], ref.act[name]), f"activation {name}"
    small = {n.name for n in net.nodes if net.route.get(n.name) == "small"}
    for name in net.vars.slots:
        a, b = net.vars.grad(name), ref.vars.grad(name)
        if name.split("/")[0] in small:      # the CUDA-core 1x1 head sums its weight gradient with fp32 atomics: order-dependent last bits
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4, f"gradient {name}"       # (relative to the tensor's max)
        else:
            assert torch.equal(a, b), f"gradient {name}"


def test_unet_training_steps(cuda_device):
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    net, variables, x, lab = _build(cuda_device, "he")
    net.keep_prepool = False      # the default training path
    step = AdamOptimizer(1e-4).minimize(net)
    orc = UNetOracle(variables, bf16_storage=True)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    got, ref = [], []
    for _ in range(5):
        got.append(float(step({net.image: xd, net.annotation: ld})))
        ref.append(orc.train_step(x, lab)[0])
    print("unet loss gpu", got, "ref", ref)
    np.testing.assert_allclose(got, ref, rtol=5e-2)
    assert got[-1] < got[0]


# ---- the reference's SegNet (SegNet.py:28-87): no ReLU, no skips, 3x3 conv to num_classes + BN -----------

def _build_segnet(cuda_device):
    from semanticsegmentation_tensorflow_b200.graph import SegNet
    variables = unet_init(3, 2, seed=1234, init="he", model="segnet")
    rng = np.random.default_rng(6)
    for k in variables:
        if k.endswith("weights"):            # the net is linear (relu=False): unit-gain init instead of He's sqrt(2)
            variables[k] = (variables[k] / np.float32(math.sqrt(2.0))).astype(np.float32)
        if k.endswith("gamma"):
            variables[k] = (1 + 0.1 * rng.standard_normal(variables[k].shape)).astype(np.float32)
        if k.endswith("beta"):
            variables[k] = (0.05 * rng.standard_normal(variables[k].shape)).astype(np.float32)
    x, lab = synthetic_batch(N, H, W, seed=0, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    net = SegNet(torch.as_tensor(x).to(cuda_device), 2, variables=variables)
    net.keep_prepool = True
    return net, variables, x, lab


def test_segnet_graph_matches_oracle_definition():
    from semanticsegmentation_tensorflow_b200.graph import graph_variable_shapes, segnet_nodes
    from oracle.graph_oracle import unet_variable_shapes
    shapes = graph_variable_shapes(segnet_nodes(2), 3)
    assert list(shapes.items()) == list(unet_variable_shapes(3, 2, "segnet").items())
    assert sum(int(np.prod(v)) for v in shapes.values()) == 39_201_092          # 39.2 M params


def test_segnet_forward_and_gradients(cuda_device):
    net, variables, x, lab = _build_segnet(cuda_device)
    pred, logits = net.create()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = UNetOracle(variables, bf16_storage=True, bf16_grads=True, model="segnet")
    loss_ref, logits_ref, grads_ref = orc.loss_and_grads(x, lab)
    _, _, grads_f32 = UNetOracle(variables, bf16_storage=False, model="segnet").loss_and_grads(x, lab)
    for name, t in net.act.items():
        e = rel_err(t.float().cpu().numpy(), orc.acts[name].detach().numpy())
        assert e <= 2e-2, f"segnet activation {name}: rel err {e:.3e}"
    assert abs(float(loss) - loss_ref) <= 2e-3 * abs(loss_ref)
    for name in net.vars.slots:
        g, r, f = net.vars.grad(name).cpu().numpy(), grads_ref[name].numpy(), grads_f32[name].numpy()
        e, c = rel_err(g, r), cosine(g, r)
        tol_e, tol_c = max(5e-2, 3 * rel_err(r, f)), min(0.999, 1 - 9 * (1 - cosine(r, f)))
        tol_e, tol_c = min(tol_e, 0.15), max(tol_c, 0.99)            # hard floor
        assert e <= tol_e and c >= tol_c, f"segnet grad {name}: rel {e:.3e} (tol {tol_e:.3e}) cos {c:.6f} (tol {tol_c:.6f})"


def test_segnet_head_on_tensor_cores_matches_cuda_core_head(cuda_device):
    """The 3x3 conv to num_classes (SegNet.py:80) as a 64-column tensor-core tile with the narrow fp32 epilogue
    (route "head_tc", taken for maps >= 4096 pixels) against the CUDA-core head kernels on the same variables:
    logits, loss and every gradient."""
    from semanticsegmentation_tensorflow_b200.graph import SegNet
    net0, variables, x0, lab0 = _build_segnet(cuda_device)
    x, lab = synthetic_batch(1, 64, 96, seed=3, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    outs = {}
    for tc in (True, False):
        net = SegNet(xd, 2, variables=variables, head_on_tensor_cores=tc)
        last = net.nodes[-1].name
        assert net.route[last] == ("head_tc" if tc else "small")
        _, logits = net.create()
        loss = net.loss(ld, with_grad=True)
        net.backward()
        net.side.join(); net.wside.join()
        torch.cuda.synchronize()
        outs[tc] = (logits.clone(), float(loss), {k: net.vars.grad(k).clone() for k in net.vars.slots})
    # (the tensor-core head multiplies bf16-rounded weights, the CUDA-core head fp32 ones: agreement at the bf16 level)
    assert rel_err(outs[True][0].cpu().numpy(), outs[False][0].cpu().numpy()) <= 5e-3
    assert abs(outs[True][1] - outs[False][1]) <= 1e-3 * abs(outs[False][1])
    for k in outs[True][2]:
        g, r = outs[True][2][k].cpu().numpy(), outs[False][2][k].cpu().numpy()
        assert rel_err(g, r) <= 2e-2 and cosine(g, r) >= 0.9995, (k, rel_err(g, r), cosine(g, r))


def test_segnet_training_steps(cuda_device):
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    net, variables, x, lab = _build_segnet(cuda_device)
    net.keep_prepool = False      # the default training path: pre-pool conv outputs are not stored
    step = AdamOptimizer(1e-4).minimize(net)
    orc = UNetOracle(variables, bf16_storage=True, model="segnet")
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    got, ref = [], []
    for _ in range(4):
        got.append(float(step({net.image: xd, net.annotation: ld})))
        ref.append(orc.train_step(x, lab)[0])
    print("segnet loss gpu", got, "ref", ref)
    np.testing.assert_allclose(got, ref, rtol=5e-2)
