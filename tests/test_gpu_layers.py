"""GPU: the reference-signature layer helpers (FCN.py:117-171) and the feed handling of the builders.

* FCN-8s composed call by call from conv_layer / max_pool / dropout / deconv_layer / fuse exactly as
  FCN.create() does in the reference (FCN.py:52-107) gives the logits of `FCN(...).create()`.
* feeds: float32 images take the bf16 first-layer path (nothing is rounded or clamped), one-hot / [N,H,W,1]
  annotations, out-of-range labels are ignored, num_classes = 4 runs through the generic loss / argmax kernels."""
import numpy as np
import pytest
import torch

from oracle.fcn_oracle import FCN8sOracle, init_variables, synthetic_batch
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu

FC = 128
N, H, W = 2, 64, 96


def _net(cuda_device, ncls=2, init="he", x=None):
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    variables = init_variables(cin=3, ncls=ncls, fc=FC, seed=1234, init=init)
    xs, lab = synthetic_batch(N, H, W, seed=0)
    xs = (xs // 32).astype(np.uint8)
    net = FCN(torch.as_tensor(xs if x is None else x).to(cuda_device), 1.0, ncls, variables=variables, fc=FC)
    return net, variables, xs, lab


def test_fcn8s_composed_from_reference_helpers(cuda_device):
    import semanticsegmentation_tensorflow_b200 as pkg
    from semanticsegmentation_tensorflow_b200.layers import VariableStore
    net, variables, x, _ = _net(cuda_device)
    _, logits = net.create()
    st = VariableStore(cuda_device)
    st.assign(variables)
    conv = lambda t, f, name, k=3: pkg.conv_layer(t, f, name, k, k, store=st)
    pool = lambda t, name: pkg.max_pool(t, name, store=st)
    h = torch.as_tensor(x).to(cuda_device)
    h = conv(h, 64, "conv1_1"); h = conv(h, 64, "conv1_2"); h = pool(h, "pool1")                 # FCN.py:52-54
    h = conv(h, 128, "conv2_1"); h = conv(h, 128, "conv2_2"); h = pool(h, "pool2")               # :56-58
    h = conv(h, 256, "conv3_1"); h = conv(h, 256, "conv3_2"); h = conv(h, 256, "conv3_3")
    pool3 = pool(h, "pool3")                                                                        # :60-63
    h = conv(pool3, 512, "conv4_1"); h = conv(h, 512, "conv4_2"); h = conv(h, 512, "conv4_3"); h = conv(h, 512, "conv4_4")
    pool4 = pool(h, "pool4")                                                                        # :65-69
    h = conv(pool4, 512, "conv5_1"); h = conv(h, 512, "conv5_2"); h = conv(h, 512, "conv5_3")
    h = pool(h, "pool5")                                                                            # :71-75
    h = pkg.dropout(conv(h, FC, "conv6", 7), 1.0, store=st)                                         # :78-79
    h = pkg.dropout(conv(h, FC, "conv7", 1), 1.0, store=st)                                         # :82-83
    h = conv(h, 2, "conv8", 1)                                                                      # :86
    t1 = pkg.deconv_layer(h, pool4.shape, 2, "conv_t1", pool4.shape, store=st)                      # :90-91
    f1 = pkg.fuse(t1, pool4, "fuse_1", store=st)                                                    # :92
    t2 = pkg.deconv_layer(f1, pool3.shape, 512, "conv_t2", pool3.shape, store=st)                   # :94-95
    f2 = pkg.fuse(t2, pool3, "fuse_2", store=st)                                                    # :96
    torch.cuda.synchronize()
    assert rel_err(f1.float().cpu().numpy(), net.act["conv_t1"].float().cpu().numpy()) <= 1e-2     # fused vs separate add
    assert rel_err(f2.float().cpu().numpy(), net.act["conv_t2"].float().cpu().numpy()) <= 1e-2
    # conv_t3 is inline in the reference (variable 'bias', 16x16 s8, FCN.py:98-107): same kernels through Ops
    e = 16 * 16 * 2
    wk, _ = st.ops.pack_matrix(st.vars["conv_t3/weights"].view(1, e, 256))
    yp = st.ops.conv2d_fwd(f2, wk, None, torch.empty((N, H // 8, W // 8, e), dtype=torch.float32, device=cuda_device), 1, 1, relu=False)
    lg = torch.empty((N, H, W, 2), dtype=torch.float32, device=cuda_device)
    st.ops.deconv_col2im(yp, st.vars["conv_t3/bias"], lg, 16, 8)
    torch.cuda.synchronize()
    assert rel_err(lg.cpu().numpy(), logits.cpu().numpy()) <= 1e-2
    # helper-level errors are loud, not fallbacks
    with pytest.raises(ValueError):
        pkg.conv_layer(h, 64, "bad", stride=2, store=st)
    with pytest.raises(ValueError):
        pkg.max_pool(h, "bad", 3, 3, 2, store=st)
    # dropout at keep_prob < 1 rescales the kept elements by 1/keep_prob
    d = pkg.dropout(pool4, 0.5, store=st, seed=7).float()
    kept = d != 0
    assert torch.equal(d[kept], (pool4.float() * 2.0).to(torch.bfloat16).float()[kept])
    nz = pool4.float() != 0
    assert 0.4 < float((kept & nz).sum()) / float(nz.sum()) < 0.6


def test_float_image_feed_takes_the_bf16_path(cuda_device):
    net, variables, x, lab = _net(cuda_device)
    _, lg_u8 = net.create()
    lg_u8 = lg_u8.clone()
    # the same pixels as float32 (the reference's placeholder dtype, FCN.py:312): integers are exact in bf16
    net.feed({net.image: torch.as_tensor(x.astype(np.float32))})
    assert net.x.dtype == torch.bfloat16
    _, lg_f = net.create()
    torch.cuda.synchronize()
    assert rel_err(lg_f.cpu().numpy(), lg_u8.cpu().numpy()) <= 1e-5
    # a mean-subtracted feed is NOT rounded / clamped to u8: it matches the oracle on the same (bf16-grid) values
    xf = (x.astype(np.float32) - 3.5) / 2.0
    xq = torch.as_tensor(xf).to(torch.bfloat16).float().numpy()
    net.feed({net.image: torch.as_tensor(xf)})
    _, lg = net.create()
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, bf16_storage=True)
    _, ref = orc.forward(xq)
    assert rel_err(lg.cpu().numpy(), ref.detach().numpy()) <= 2e-2
    with pytest.raises(TypeError):
        net.feed({net.image: torch.as_tensor(x.astype(np.float64))})


def test_annotation_forms_and_ignored_labels(cuda_device):
    net, variables, x, lab = _net(cuda_device)
    net.forward()
    l_ids = float(net.loss(torch.as_tensor(lab)))
    onehot = np.stack([lab == 0, lab == 1], axis=-1)                       # process_gt_image, FCN.py:195-201 (bool)
    net.forward()
    l_bool = float(net.loss(torch.as_tensor(onehot)))
    net.forward()
    l_f32 = float(net.loss(torch.as_tensor(onehot.astype(np.float32)).to(cuda_device)))
    net.forward()
    l_n1 = float(net.loss(torch.as_tensor(lab[..., None].astype(np.int64))))     # [N,H,W,1] ids (FCN.py:329 variant)
    assert l_ids == l_bool == l_f32 == l_n1
    with pytest.raises(ValueError):
        net.loss(torch.zeros((N, H, W, 3)))
    # labels >= num_classes (e.g. a 255 "ignore" value) contribute no loss and no gradient
    lab2 = lab.copy()
    lab2[:, : H // 2] = 255
    net.forward()
    l_ign = float(net.loss(torch.as_tensor(lab2), with_grad=True))
    torch.cuda.synchronize()
    assert float(net.dlogits[:, : H // 2].abs().max()) == 0.0
    lg = net.logits.double().cpu()
    per_px = torch.nn.functional.cross_entropy(lg.view(-1, 2), torch.as_tensor(lab).long().view(-1), reduction="none").view(N, H, W)
    assert abs(l_ign - float(per_px[:, H // 2:].sum() / (N * H * W))) <= 1e-5 * abs(l_ign)
    cm = net.confusion_matrix()
    assert int(cm.sum()) == N * H * W       # the standalone confusion kernel counts label & 1 (documented: ids only)


def test_four_classes_generic_loss_and_argmax(cuda_device):
    net, variables, x, _ = _net(cuda_device, ncls=4)
    rng = np.random.default_rng(5)
    lab = rng.integers(0, 4, (N, H, W), dtype=np.uint8)
    pred, logits = net.create()
    loss = float(net.loss(torch.as_tensor(lab), with_grad=True))
    torch.cuda.synchronize()
    orc = FCN8sOracle(variables, ncls=4, bf16_storage=True)
    pred_ref, logits_ref = orc.forward(x)
    assert pred.shape == (N, H, W, 1) and pred.dtype == torch.int64
    assert rel_err(logits.cpu().numpy(), logits_ref.detach().numpy()) <= 2e-2
    lg = logits.cpu()
    assert torch.equal(pred.cpu()[..., 0], lg.argmax(-1))                    # first index on ties == torch.argmax
    ref_loss = float(orc.loss(logits_ref, lab))
    assert abs(loss - ref_loss) <= 2e-2 * abs(ref_loss)
    oh = torch.nn.functional.one_hot(torch.as_tensor(lab).long(), 4)
    d_ref = (torch.softmax(lg.double(), -1) - oh) / (N * H * W)
    assert rel_err(net.dlogits.cpu().numpy(), d_ref.numpy()) <= 1e-5
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    from semanticsegmentation_tensorflow_b200._lib import SegkError
    with pytest.raises((SegkError, ValueError)):
        FCN(torch.as_tensor(x).to(cuda_device), 1.0, 3, fc=FC).create()        # 3 classes: no kernel, no fallback
