"""FCDenseNet (FCDenseNet.py:23-163).  CPU: the product's graph / variable inventory against the independent oracle
restatement and the SURVEY's numbers, the physical channel layouts.  GPU: activations, every gradient and Adam steps
against the oracle on a reduced configuration whose channel counts hit the padding cases of the real one
(segments that are not multiples of 8: 44 + 16 j, 46 + 16 j, like 140 and 174)."""
import numpy as np
import pytest
import torch

from oracle.densenet_oracle import FCDenseNetOracle, densenet_init
from oracle.fcn_oracle import synthetic_batch
from tests.gpu_util import cosine, rel_err

SMALL = dict(n_layers_per_blocks=(1, 2, 2), growth_rate=16, n_filters_first_conv=40, theta=0.5)
N, H, W = 2, 32, 48


def test_graph_and_variables_match_the_oracle_and_the_survey():
    from semanticsegmentation_tensorflow_b200.densenet import (densenet_variable_shapes, fcdensenet_flops_per_image,
                                                               fcdensenet_nodes)
    nodes, ch = fcdensenet_nodes(2)
    shapes = densenet_variable_shapes(nodes, ch)
    assert list(shapes.items()) == list(FCDenseNetOracle(None, 2).variable_shapes().items())
    assert sum(1 for n in nodes if n.kind in ("conv", "deconv")) == 130
    assert sum(int(np.prod(s)) for s in shapes.values()) == 10_579_976                      # "10.6 M params"
    finals = [ch[n.name] for n in nodes if n.kind == "concat" and n.name.endswith("final")]
    assert finals == [128, 160, 208, 280, 348, 430]                                         # dense-block widths
    assert [ch[n.inputs[0]] for n in nodes if n.kind == "deconv"] == [430, 696, 560, 416, 320]   # FCDenseNet.py:141-153
    fwd, train = fcdensenet_flops_per_image(nodes, ch, 160, 576)
    assert 70e9 < fwd < 74e9                                                                # 73.6 GFLOP dense-tap
    nodes_s, ch_s = fcdensenet_nodes(2, **SMALL)
    assert list(densenet_variable_shapes(nodes_s, ch_s).items()) == list(FCDenseNetOracle(None, 2, **SMALL).variable_shapes().items())
    assert any(c % 8 for c in ch_s.values() if c > 8)         # the reduced net exercises the non-multiple-of-8 cases


def test_physical_layout_pads_segments_to_8_and_total_to_64():
    from semanticsegmentation_tensorflow_b200.densenet import _Layout
    L = _Layout([140, 16, 16])
    assert (L.logical, L.used, L.cp) == (172, 176, 192)
    assert list(L.cmap[:140]) == list(range(140)) and list(L.cmap[140:144]) == [-1] * 4
    assert list(L.cmap[144:176]) == list(range(140, 172)) and set(L.cmap[176:]) == {-1}
    assert sorted(c for c in L.cmap if c >= 0) == list(range(172))


def _build(cuda_device, init="he", keep=1.0):
    from semanticsegmentation_tensorflow_b200.densenet import FCDenseNet
    orc0 = FCDenseNetOracle(None, 2, **SMALL)
    variables = densenet_init(orc0.variable_shapes(), seed=1234, init=init)
    rng = np.random.default_rng(3)
    for k in variables:                                  # non-trivial BN parameters
        if k.endswith("gamma"):
            variables[k] = (1 + 0.2 * rng.standard_normal(variables[k].shape)).astype(np.float32)
        if k.endswith("beta"):
            variables[k] = (0.1 * rng.standard_normal(variables[k].shape)).astype(np.float32)
    x, lab = synthetic_batch(N, H, W, seed=0, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    net = FCDenseNet(torch.as_tensor(x).to(cuda_device), keep, 2, variables=variables, **SMALL)
    return net, variables, x, lab


@pytest.mark.gpu
def test_fcdensenet_forward_and_all_gradients(cuda_device):
    net, variables, x, lab = _build(cuda_device)
    pred, logits = net.create()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = FCDenseNetOracle(variables, 2, bf16_storage=True, bf16_grads=True, **SMALL)
    loss_ref, logits_ref, grads_ref = orc.loss_and_grads(x, lab)
    _, _, grads_f32 = FCDenseNetOracle(variables, 2, **SMALL).loss_and_grads(x, lab)
    # conv / deconv outputs (logical channels of the padded tensors) and the dense-block buffers
    for name, ref in orc.acts.items():
        ref = ref.detach().numpy()
        if name.startswith("denseblock") and name in [f"denseblock{b}" for b in range(1, 10)]:
            root = f"{name}bottleneck_layer_concatenate_final"
            L = net.lay[root]
            got = net.buf[root].float().cpu().numpy()[..., L.cmap >= 0]
        elif name == "final_conv":
            got = net.logits.cpu().numpy()
        else:
            got = net.act[name].float().cpu().numpy()[..., :ref.shape[3]]
        e = rel_err(got, ref)
        assert e <= 2e-2, f"activation {name}: rel err {e:.3e}"
    assert pred.shape == (N, H, W, 1) and pred.dtype == torch.int64
    assert rel_err(logits.cpu().numpy(), logits_ref.numpy()) <= 2e-2
    assert abs(float(loss) - loss_ref) <= 2e-3 * abs(loss_ref)
    # pads of every stored tensor are exactly zero
    for name, t in net.buf.items():
        assert float(t[..., torch.as_tensor(net.lay[name].cmap < 0)].abs().max()) == 0.0, name
    worst = ("", 0.0, 1.0)
    for name in net.vars.slots:
        g, r, f = net.vars.grad(name).cpu().numpy(), grads_ref[name].numpy(), grads_f32[name].numpy()
        e, c = rel_err(g, r), cosine(g, r)
        tol_e, tol_c = max(5e-2, 3 * rel_err(r, f)), min(0.999, 1 - 9 * (1 - cosine(r, f)))
        tol_e, tol_c = min(tol_e, 0.15), max(tol_c, 0.99)
        assert e <= tol_e and c >= tol_c, f"grad {name}: rel {e:.3e} (tol {tol_e:.3e}) cos {c:.6f} (tol {tol_c:.6f})"
        if e > worst[1]:
            worst = (name, e, c)
    print(f"fcdensenet worst grad {worst[0]}: rel {worst[1]:.3e} cos {worst[2]:.6f}")


@pytest.mark.gpu
def test_fcdensenet_dropout_with_injected_masks(cuda_device):
    """keep_prob 0.8 with the same keep masks on both sides (TF's RNG stream is not reproducible, FCDenseNet.py:30,34)."""
    net, variables, x, lab = _build(cuda_device, keep=0.8)
    rng = np.random.default_rng(11)
    masks_ref, masks_dev = {}, {}
    for n in net.nodes:
        if n.kind != "dropout":
            continue
        conv = net._base(n.name)
        h, w = net.hw[conv]
        c, cp = net.ch[conv], net.lay[conv].cp
        m = (rng.random((N, h, w, c)) < 0.8).astype(np.float32)
        masks_ref[conv] = torch.tensor(m)
        mp = np.ones((N, h, w, cp), np.uint8)
        mp[..., :c] = m.astype(np.uint8)
        masks_dev[n.name] = torch.as_tensor(mp).to(cuda_device)
    net.injected_masks = masks_dev
    net.forward()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = FCDenseNetOracle(variables, 2, bf16_storage=True, bf16_grads=True, **SMALL)
    loss_ref, logits_ref, grads_ref = orc.loss_and_grads(x, lab, keep_prob=0.8, masks=masks_ref)
    assert rel_err(net.logits.cpu().numpy(), logits_ref.numpy()) <= 2e-2
    assert abs(float(loss) - loss_ref) <= 5e-3 * abs(loss_ref)
    for name in ("dense_init/weights", "denseblock2bottleneck_layer_1_conv2/weights", "transition_up1/weights",
                 "batch_normalization_3/gamma", "final_conv/weights"):
        g, r = net.vars.grad(name).cpu().numpy(), grads_ref[name].numpy()
        assert rel_err(g, r) <= 0.1 and cosine(g, r) >= 0.99, (name, rel_err(g, r), cosine(g, r))
    # Philox masks: the keep rate is right and the step runs
    net.injected_masks = None
    net.forward()
    # (the Dropout nodes are applied on the fly by their readers: the dropped conv2 output is visible in its concat slot)
    r, off = net.member["denseblock1bottleneck_layer_0_conv2"]
    t = net.buf[r][..., off:off + 16].float()
    assert 0.7 < float((t != 0).float().mean()) <= 0.85
    assert set(net.drop_fused.values()) == {"bn", "copy"} and len(net.drop_fused) == sum(n.kind == "dropout" for n in net.nodes)


@pytest.mark.gpu
def test_fcdensenet_training_steps(cuda_device):
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    net, variables, x, lab = _build(cuda_device)
    step = AdamOptimizer(1e-4).minimize(net)
    orc = FCDenseNetOracle(variables, 2, bf16_storage=True, **SMALL)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    got, ref = [], []
    for _ in range(5):
        got.append(float(step({net.image: xd, net.annotation: ld, net.keep_probability: 1.0})))
        ref.append(orc.train_step(x, lab)[0])
    print("fcdensenet loss gpu", got, "ref", ref)
    np.testing.assert_allclose(got, ref, rtol=5e-2)
    assert got[-1] < got[0]
    for name in ("dense_init/weights", "batch_normalization_5/gamma", "transition_up2/weights"):
        assert rel_err(net.vars.param(name).cpu().numpy(), orc.vars[name].detach().numpy()) <= 2e-2, name


@pytest.mark.gpu
def test_fcdensenet_full_configuration_runs(cuda_device):
    """The reference's own configuration (4,5,7,10,12,15 layers, 130 convs) at 160x576, batch 1: forward + backward +
    Adam, loss == ln 2 under the reference init, finite gradients, logits against the oracle."""
    import math
    from semanticsegmentation_tensorflow_b200.densenet import FCDenseNet
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    variables = densenet_init(FCDenseNetOracle(None, 2).variable_shapes(), seed=1234, init="he")
    x, lab = synthetic_batch(1, 160, 576, seed=0, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    net = FCDenseNet(torch.as_tensor(x).to(cuda_device), 1.0, 2, variables=variables)
    _, logits = net.create()
    torch.cuda.synchronize()
    _, logits_ref = FCDenseNetOracle(variables, 2, bf16_storage=True).forward(x)
    assert rel_err(logits.cpu().numpy(), logits_ref.detach().numpy()) <= 3e-2
    step = AdamOptimizer(1e-4).minimize(net)
    l0 = float(step({net.image: torch.as_tensor(x).to(cuda_device), net.annotation: torch.as_tensor(lab).to(cuda_device)}))
    assert math.isfinite(l0) and bool(torch.isfinite(net.vars.g).all()) and float(net.vars.g.abs().max()) > 0
