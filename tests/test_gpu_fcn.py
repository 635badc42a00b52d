"""GPU parity: the whole FCN-8s graph (forward, every saved activation, every one of the 40
gradient tensors, Adam steps) through the reference-shaped API vs the CPU oracle that mirrors
the bf16 storage points.  Two inits (SURVEY §0 finding 5): 'he' makes every layer numerically
visible; 'ref' is the reference's N(0, 0.01^2), where the initial loss is ln 2.

Tolerances (north_star): logits rtol 2e-2 of the tensor max under bf16; argmax agreement >= 99.9 %
on pixels whose logit margin exceeds 1 % of mean|logit|.  Gradients: 5e-2 of the tensor max and
cosine >= 0.999, widened per tensor to 3x the discrepancy between the bf16-mirroring oracle and the
plain fp32 oracle when that is larger (ReLU-mask flips at near-zero activations make a few deep,
small-spatial gradients noisier than rounding alone)."""
import math

import numpy as np
import pytest
import torch

from oracle.fcn_oracle import FCN8sOracle, init_variables, synthetic_batch
from tests.gpu_util import cosine, rel_err

pytestmark = pytest.mark.gpu

FC = 128
N, H, W = 2, 128, 192
# hard floor under the adaptive gradient tolerances: whatever the oracle's own bf16-vs-fp32 noise is,
# a gradient tensor below cosine 0.99 or beyond 15 % of its max fails
HARD_COS, HARD_REL = 0.99, 0.15


def _build(cuda_device, init, keep=1.0, scale_input=True):
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    variables = init_variables(cin=3, ncls=2, fc=FC, seed=1234, init=init)
    x, lab = synthetic_batch(N, H, W, seed=0, road_shaped=True)
    if scale_input and init == "he":
        x = (x // 32).astype(np.uint8)       # O(1) activations under He init
    net = FCN(torch.as_tensor(x).to(cuda_device), keep, 2, variables=variables, fc=FC)
    return net, variables, x, lab


@pytest.mark.parametrize("init", ["he", "ref"])
def test_forward_activations_and_logits(cuda_device, init):
    net, variables, x, lab = _build(cuda_device, init)
    net.keep_prepool = True          # also store the pre-pool conv outputs (default: the fused pool writes only the pooled tensor)
    pred, logits = net.create()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True)
    pred_ref, logits_ref = orc.forward(x)
    names = {"conv_t1": "fuse_1", "conv_t2": "fuse_2", "conv_t3": "logits"}
    worst = {}
    for name, t in net.act.items():
        ref = orc.acts[names.get(name, name)].detach().numpy()
        got = t.float().cpu().numpy()
        if name == "conv8":
            # 4096->2 dot products of zero-mean weights with post-ReLU inputs cancel heavily (the bf16
            # oracle itself differs from the fp32 oracle by 3-4 % of max here): measure against the
            # un-cancelled scale |a7| . |w8|
            import oracle.tf_ops as T
            scale = T.conv2d_same(orc.acts["dropout7"].detach().abs(), orc.vars["conv8/weights"].detach().abs()).max()
            e = float(np.abs(got - ref).max() / float(scale))
        else:
            e = rel_err(got, ref)
        worst[name] = e
        assert e <= 2e-2, f"[{init}] activation {name}: rel err {e:.3e}"
    lg, lr = logits.cpu().numpy(), logits_ref.detach().numpy()
    assert pred.shape == (N, H, W, 1) and pred.dtype == torch.int64
    margin = np.abs(lr[..., 1] - lr[..., 0])
    sel = margin >= 0.01 * np.abs(lr).mean()
    agree = (pred.cpu().numpy()[..., 0] == pred_ref.numpy()[..., 0])
    assert agree[sel].mean() >= 0.999, f"[{init}] argmax agreement {agree[sel].mean():.5f} on {sel.mean():.2%} of pixels"
    print(f"[{init}] worst activation rel err {max(worst.values()):.3e}; raw argmax agreement {agree.mean():.5f}")


@pytest.mark.parametrize("init", ["he", "ref"])
def test_loss_and_all_gradients(cuda_device, init):
    net, variables, x, lab = _build(cuda_device, init)
    net.forward()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True, bf16_grads=True)
    loss_ref, _, grads_ref = orc.loss_and_grads(x, lab)
    _, _, grads_f32 = FCN8sOracle(variables, bf16_storage=False).loss_and_grads(x, lab)
    assert abs(float(loss) - loss_ref) <= 1e-3 * abs(loss_ref)
    if init == "ref":
        assert abs(float(loss) - math.log(2.0)) < 1e-3            # SURVEY §0 finding 5
    report = []
    for name in net.vars.slots:
        g = net.vars.grad(name).cpu().numpy()
        r = grads_ref[name].numpy()
        e, c = rel_err(g, r), cosine(g, r)
        f = grads_f32[name].numpy()
        e0, c0 = rel_err(r, f), cosine(r, f)                        # the oracle's own bf16 noise
        tol_e, tol_c = max(5e-2, 3 * e0), min(0.999, 1 - 9 * (1 - c0))   # 3x in amplitude = 9x in cosine deficit
        tol_e, tol_c = min(tol_e, HARD_REL), max(tol_c, HARD_COS)        # ... but never beyond the hard floor
        report.append((name, e, c, e0, c0))
        assert e <= tol_e and c >= tol_c, (f"[{init}] grad {name}: rel err {e:.3e} (tol {tol_e:.3e}) cosine {c:.6f} "
                                           f"(tol {tol_c:.6f}) max|ref| {np.abs(r).max():.3e}")
    worst = max(report, key=lambda t: t[1])
    print(f"[{init}] worst grad {worst[0]}: rel {worst[1]:.3e} cos {worst[2]:.6f} (oracle bf16-vs-fp32: rel {worst[3]:.3e} cos {worst[4]:.6f})")
    for name, e, c, e0, c0 in report:
        print(f"   {name:20s} rel {e:.3e} cos {c:.6f} | oracle self-noise rel {e0:.3e} cos {c0:.6f}")
    cm = net.confusion_matrix().cpu().numpy()
    pred_ref = orc.acts["logits"].detach().numpy().argmax(-1)
    assert cm.sum() == N * H * W


def _curves(cuda_device, lr, steps):
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    net, variables, x, lab = _build(cuda_device, "he")
    step = AdamOptimizer(lr).minimize(net)
    orc = FCN8sOracle(variables, bf16_storage=True)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    got, ref = [], []
    for _ in range(steps):
        got.append(float(step({net.image: xd, net.annotation: ld, net.keep_probability: 1.0})))
        ref.append(orc.train_step(x, lab, lr=lr)[0])
    return net, orc, x, np.array(got), np.array(ref)


def test_training_curve_100_steps_and_raw_argmax(cuda_device):
    """north_star: "a matching loss curve over 100 steps" and "argmax agreement >= 99.9 %" (raw, every pixel).
    100 TF-Adam steps (FCN.py:338-340,398) at keep_prob 1.0 under the He init, GPU path vs oracle.

    Two learning rates.  At the reference's 1e-4 this 2-image problem is fitted (loss 9.4 -> ~0.2) within ~35
    steps and then turns chaotic: both paths show Adam loss spikes, at different steps (measured: oracle 0.26 at
    step 99 after 0.007 at step 90; GPU 0.18 at step 40) -- pointwise agreement is asserted on the first 30 steps
    and the tail is only required to keep training.  At 2e-5 the descent stays smooth and all 100 points must match.
    Adam's first steps move every weight by ~lr * sign(g): bf16 noise flips the sign of near-zero gradients, so
    trajectories agree to a few per cent, not to rounding."""
    net, orc, x, got, ref = _curves(cuda_device, 2e-5, 100)
    dev = np.abs(got - ref) / ref
    print("lr 2e-5 loss curve gpu", ["%.5f" % v for v in got[::10]], "%.5f" % got[-1])
    print("lr 2e-5 loss curve ref", ["%.5f" % v for v in ref[::10]], "%.5f" % ref[-1])
    print(f"lr 2e-5 relative deviation over 100 steps: max {dev.max():.3e} mean {dev.mean():.3e}")
    assert ref[-1] < 0.5 * ref[0] and got[-1] < 0.5 * got[0]            # it trains
    assert dev.max() <= 0.10 and dev.mean() <= 3e-2, (dev.max(), dev.mean())
    # variables after 100 steps
    for name in ("conv1_1/weights", "conv5_3/weights", "conv_t3/weights", "conv8/biases"):
        p = net.vars.param(name).cpu().numpy()
        r = orc.vars[name].detach().numpy()
        assert rel_err(p, r) <= 5e-2, (name, rel_err(p, r))
    # raw argmax agreement after training: (a) each path with its own trained weights
    pred, _ = net.create()
    pred_ref, _ = orc.forward(x)
    own = float((pred.cpu().numpy() == pred_ref.numpy()).mean())
    # (b) the GPU path on the ORACLE's trained weights: isolates the forward arithmetic from the
    # trajectory drift; this is the >= 99.9 % criterion, over every pixel
    net.vars.assign({k: v.detach().numpy() for k, v in orc.vars.items()})
    net.vars.repack(net.ops)
    pred2, logits2 = net.create()
    same = pred2.cpu().numpy() == pred_ref.numpy()
    same_w = float(same.mean())
    print(f"raw argmax agreement after 100 steps: own weights {own:.5f}, same weights {same_w:.5f}")
    # After 100 steps at this small learning rate a fraction of the pixels still sits within bf16 noise of the decision
    # boundary, and the count of flipped pixels moves by one or two with the fp32 summation order of a single kernel
    # (measured 0.99906 / 0.99900 / 0.99898 = 46 / 49 / 50 of 49152 pixels under three schedules of conv5/conv6).  So:
    # >= 99.8 % raw here, EVERY disagreeing pixel must be a near-tie of the oracle (|logit margin| within the bf16
    # tolerance of the logits), and the >= 99.9 % raw criterion is asserted below on the model trained at the
    # reference's own learning rate (measured 99.99 %).
    assert same_w >= 0.998, same_w
    lr_ = orc.acts["logits"].detach().numpy()
    margin = np.abs(lr_[..., 1] - lr_[..., 0])[~same[..., 0]] if same.ndim == 4 else np.abs(lr_[..., 1] - lr_[..., 0])[~same]
    assert margin.size == 0 or margin.max() <= 2e-2 * np.abs(lr_).max(), (margin.max(), np.abs(lr_).max())
    # (a) compares two 100-step trajectories: it moves with the fp32 summation order of the kernels (measured 0.9911
    # with 2-way split-K in conv6's dgrad, 0.9891 with the lockstep tap-split schedule), so it only guards against
    # gross divergence; (b) is the acceptance criterion
    assert own >= 0.98, own
    # the reference's learning rate
    net, orc, x, got, ref = _curves(cuda_device, 1e-4, 100)
    dev = np.abs(got - ref) / ref
    print("lr 1e-4 loss curve gpu", ["%.5f" % v for v in got[::10]], "%.5f" % got[-1])
    print("lr 1e-4 loss curve ref", ["%.5f" % v for v in ref[::10]], "%.5f" % ref[-1])
    print(f"lr 1e-4 relative deviation: first 30 steps max {dev[:30].max():.3e}; all 100 max {dev.max():.3e} mean {dev.mean():.3e}")
    np.testing.assert_allclose(got[:30], ref[:30], rtol=5e-2)
    assert np.median(got[70:]) < 0.02 * got[0] and np.median(ref[70:]) < 0.02 * ref[0]     # both keep fitting
    pred_ref, _ = orc.forward(x)
    net.vars.assign({k: v.detach().numpy() for k, v in orc.vars.items()})
    net.vars.repack(net.ops)
    pred2, _ = net.create()
    same_w = float((pred2.cpu().numpy() == pred_ref.numpy()).mean())
    print(f"lr 1e-4 raw argmax agreement after 100 steps, same weights: {same_w:.5f}")
    assert same_w >= 0.999, same_w


def test_training_curve_full_size_fc4096(cuda_device):
    """The benchmarked configuration itself (fc = 4096, 160x576), batch 1: 12 Adam steps vs the oracle."""
    from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer
    variables = init_variables(cin=3, ncls=2, fc=4096, seed=1234, init="he")
    x, lab = synthetic_batch(1, 160, 576, seed=0, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    net = FCN(torch.as_tensor(x).to(cuda_device), 1.0, 2, variables=variables)
    step = AdamOptimizer(1e-4).minimize(net)
    orc = FCN8sOracle(variables, bf16_storage=True)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    got, ref = [], []
    for _ in range(12):
        got.append(float(step({net.image: xd, net.annotation: ld, net.keep_probability: 1.0})))
        ref.append(orc.train_step(x, lab)[0])
    print("full-size loss curve gpu", ["%.5f" % v for v in got])
    print("full-size loss curve ref", ["%.5f" % v for v in ref])
    assert ref[-1] < ref[0]
    np.testing.assert_allclose(got, ref, rtol=5e-2)


def test_dropout_training_step_with_injected_masks(cuda_device):
    net, variables, x, lab = _build(cuda_device, "he", keep=0.8)
    rng = np.random.default_rng(3)
    h, w = H // 32, W // 32
    masks = {k: (rng.random((N, h, w, FC)) < 0.8).astype(np.float32) for k in ("dropout6", "dropout7")}
    net.injected_masks = {k: torch.as_tensor(v.astype(np.uint8)).to(cuda_device) for k, v in masks.items()}
    net.forward()
    loss = net.loss(torch.as_tensor(lab).to(cuda_device), with_grad=True)
    net.backward()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True, bf16_grads=True)
    tmasks = {k: torch.tensor(v) for k, v in masks.items()}
    loss_ref, _, grads_ref = orc.loss_and_grads(x, lab, keep_prob=0.8, masks=tmasks)
    _, _, grads_f32 = FCN8sOracle(variables, bf16_storage=False).loss_and_grads(x, lab, keep_prob=0.8, masks=tmasks)
    assert abs(float(loss) - loss_ref) <= 2e-3 * abs(loss_ref)
    for name in ("conv6/weights", "conv7/weights", "conv5_1/weights", "conv8/weights"):
        g, r, f = net.vars.grad(name).cpu().numpy(), grads_ref[name].numpy(), grads_f32[name].numpy()
        tol_e, tol_c = max(5e-2, 3 * rel_err(r, f)), min(0.999, 1 - 9 * (1 - cosine(r, f)))
        tol_e, tol_c = min(tol_e, HARD_REL), max(tol_c, HARD_COS)
        assert rel_err(g, r) <= tol_e and cosine(g, r) >= tol_c, (name, rel_err(g, r), cosine(g, r), tol_e, tol_c)


def test_inference_softmax_and_road_mask(cuda_device):
    net, variables, x, lab = _build(cuda_device, "he")
    prob, mask = net.infer()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True)
    _, logits_ref = orc.forward(x)
    lg = net.logits.cpu()
    assert rel_err(lg.numpy(), logits_ref.detach().numpy()) <= 2e-2
    # the softmax / mask kernel itself, on the logits it was given (He-init logits saturate the
    # softmax, so probabilities are compared on identical logits)
    p_ref = torch.softmax(lg, dim=-1).numpy()
    assert np.abs(prob.cpu().numpy() - p_ref).max() <= 1e-6
    assert np.array_equal(mask.cpu().numpy(), (lg[..., 1] > lg[..., 0]).numpy().astype(np.uint8))
    assert mask.shape == (N, H, W)


def test_graphed_inference_matches_eager(cuda_device):
    """net.infer_graphed: the forward + softmax + road mask replayed from a CUDA graph give the bits of net.infer, for
    the captured image and for later images (host and device feeds)."""
    net, variables, x, lab = _build(cuda_device, "he")
    rng = np.random.default_rng(5)
    imgs = [x, rng.integers(0, 8, x.shape, dtype=np.uint8), rng.integers(0, 8, x.shape, dtype=np.uint8)]
    for i, im in enumerate(imgs):
        t = torch.as_tensor(im)
        feed = t.pin_memory() if i % 2 else t.to(cuda_device)
        p0, m0 = net.infer(t.to(cuda_device))
        p0, m0 = p0.clone(), m0.clone()
        p1, m1 = net.infer_graphed(feed)
        torch.cuda.synchronize()
        assert torch.equal(p1, p0) and torch.equal(m1, m0), f"image {i}"
    # training still works on the same net afterwards (the capture left no state behind)
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    step = AdamOptimizer(1e-4).minimize(net)
    l0 = float(step({net.image: torch.as_tensor(x).to(cuda_device), net.annotation: torch.as_tensor(lab).to(cuda_device)}))
    assert np.isfinite(l0)


def test_full_resolution_inference_384x1248(cuda_device):
    """BASELINE configs[3] shape: FCN-8s forward at 384x1248 with the real fc=4096 head, batch 1,
    vs the fp32-arithmetic oracle mirroring bf16 storage (logits rtol 2e-2, margin-conditioned argmax
    agreement >= 99.9 %), plus the softmax / road-mask epilogue (FCN.py:229,204-206)."""
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    n, h, w = 1, 384, 1248
    variables = init_variables(cin=3, ncls=2, fc=4096, seed=1234, init="ref")
    x, _ = synthetic_batch(n, h, w, seed=7)
    net = FCN(torch.as_tensor(x).to(cuda_device), 1.0, 2, variables=variables)
    prob, mask = net.infer()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True)
    pred_ref, logits_ref = orc.forward(x)
    lr = logits_ref.detach().numpy()
    lg = net.logits.cpu().numpy()
    assert rel_err(lg, lr) <= 2e-2
    margin = np.abs(lr[..., 1] - lr[..., 0])
    sel = margin >= 0.01 * np.abs(lr).mean()
    agree = mask.cpu().numpy() == pred_ref.numpy()[..., 0]
    assert agree[sel].mean() >= 0.999, agree[sel].mean()
    assert abs(float(prob.sum()) - n * h * w) <= 1e-3 * n * h * w        # softmax rows sum to 1


def test_checkpoint_roundtrip_with_adam_slots(cuda_device, tmp_path):
    """Save after 2 steps, restore into a fresh net + optimizer, and the third step is bit-identical
    (reference variable names / layouts, Adam slots, beta powers)."""
    from semanticsegmentation_tensorflow_b200.checkpoint import load_checkpoint, save_checkpoint, state_dict
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer, FCN
    net, variables, x, lab = _build(cuda_device, "he")
    net.side.enabled = net.wside.enabled = False  # deterministic kernel order for the bit-identity check
    opt = AdamOptimizer(1e-4)
    step = opt.minimize(net)
    xd, ld = torch.as_tensor(x).to(cuda_device), torch.as_tensor(lab).to(cuda_device)
    for _ in range(2):
        step({net.image: xd, net.annotation: ld})
    path = str(tmp_path / "fcn.npz")
    save_checkpoint(path, net, opt)
    sd = state_dict(net, opt)
    assert sd["conv1_1/weights"].shape == (3, 3, 3, 64) and sd["conv_t2/weights"].shape == (4, 4, 256, 512)
    assert "conv_t3/bias" in sd and "conv6/weights/Adam_1" in sd
    assert float(sd["beta1_power"]) == pytest.approx(0.9 ** 3)
    net2 = FCN(xd, 1.0, 2, variables=None, init="ref", fc=FC)
    net2.side.enabled = net2.wside.enabled = False
    opt2 = AdamOptimizer(1e-4)
    step2 = opt2.minimize(net2)
    load_checkpoint(path, net2, opt2)
    assert opt2.t == 2
    torch.cuda.synchronize()
    # the restored state is bit-identical: parameters, both Adam slots, and the repacked bf16 shadows
    assert torch.equal(net.vars.p, net2.vars.p) and torch.equal(net.vars.m, net2.vars.m) and torch.equal(net.vars.v, net2.vars.v)
    for k in net.vars.wk:
        assert torch.equal(net.vars.wk[k], net2.vars.wk[k]), k
    # and training continues from it: same loss on the next step (forward is deterministic up to the
    # fp32 atomics of the split-K layers of this small net)
    l3 = float(step({net.image: xd, net.annotation: ld}))
    l3b = float(step2({net2.image: xd, net2.annotation: ld}))
    assert abs(l3 - l3b) <= 1e-3 * abs(l3)
