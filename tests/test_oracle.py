"""Oracle self-checks (CPU): the torch restatement vs naive NumPy loops, fp64 finite
differences, TF-Adam closed form, and the ln2 initial loss (SURVEY §4 tier 1)."""
import math

import numpy as np
import pytest
import torch

from oracle import naive, tf_ops as T
from oracle.fcn_oracle import FCN8sOracle, init_variables, synthetic_batch, variable_shapes


@pytest.mark.parametrize("k,s,h,w", [(3, 1, 5, 7), (7, 1, 5, 6), (1, 1, 4, 4), (3, 2, 7, 8), (4, 2, 6, 6)])
def test_conv2d_same_vs_naive(k, s, h, w):
    rng = np.random.default_rng(k * 10 + s)
    x = rng.standard_normal((2, h, w, 3)).astype(np.float32)
    wt = rng.standard_normal((k, k, 3, 4)).astype(np.float32)
    got = T.conv2d_same(torch.tensor(x), torch.tensor(wt), s).numpy()
    ref = naive.conv2d_same_naive(x, wt, s)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k,s", [(4, 2), (16, 8)])
def test_conv2d_transpose_vs_naive_and_grad_definition(k, s):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((1, 3, 4, 2)).astype(np.float32)
    wt = rng.standard_normal((k, k, 3, 2)).astype(np.float32)   # [kh,kw,Cout,Cin]
    oh, ow = 3 * s, 4 * s
    got = T.conv2d_transpose_same(torch.tensor(x), torch.tensor(wt), (oh, ow), s).numpy()
    ref = naive.conv2d_transpose_same_naive(x, wt, (oh, ow), s)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
    # definition: input-gradient of the SAME conv2d whose input has output_shape
    inp = torch.zeros((1, oh, ow, 3), dtype=torch.float64, requires_grad=True)
    out = T.conv2d_same(inp, torch.tensor(wt, dtype=torch.float64), s)
    out.backward(torch.tensor(x, dtype=torch.float64))
    np.testing.assert_allclose(inp.grad.numpy(), ref, rtol=1e-9, atol=1e-9)


def test_max_pool_argmax_first_max_and_grad():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 3, (2, 6, 8, 5)).astype(np.float32)    # many ties
    y, idx = T.max_pool_2x2_with_argmax(x)
    yn, idxn = naive.max_pool_2x2_naive(x)
    assert np.array_equal(y, yn) and np.array_equal(idx, idxn)
    yt = T.max_pool_2x2(torch.tensor(x)).numpy()
    assert np.array_equal(y, yt)
    # gradient routing equals torch autograd (== TF MaxPoolGrad first-max rule)
    xt = torch.tensor(x, requires_grad=True)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    T.max_pool_2x2(xt).backward(torch.tensor(dy))
    assert np.array_equal(T.max_pool_2x2_grad(dy, idx, (6, 8)), xt.grad.numpy())


def test_softmax_xent_and_grad():
    rng = np.random.default_rng(1)
    lg = rng.standard_normal((2, 3, 4, 2)).astype(np.float32) * 5
    lab = rng.integers(0, 2, (2, 3, 4))
    onehot = np.eye(2, dtype=np.float32)[lab]
    lt = torch.tensor(lg, requires_grad=True)
    loss = T.softmax_cross_entropy_with_logits(lt, torch.tensor(onehot))
    ref_loss, ref_grad = naive.softmax_xent_naive(lg, onehot)
    np.testing.assert_allclose(loss.detach().numpy(), ref_loss, rtol=1e-5, atol=1e-6)
    loss.sum().backward()
    np.testing.assert_allclose(lt.grad.numpy(), ref_grad, rtol=1e-5, atol=1e-6)


def test_adam_tf_formula_closed_form_and_differs_from_torch():
    g = 1e-10
    p, m, v = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    T.adam_tf_step(p, m, v, torch.full((1,), g), 1)
    lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = -lr_t * (0.1 * g) / (math.sqrt(0.001 * g * g) + 1e-8)
    assert abs(float(p) - exp) <= 1e-6 * abs(exp)
    assert abs(float(p)) == pytest.approx(3.16e-8, rel=2e-2)
    # three steps vs explicit recurrence in fp64
    p, m, v = torch.zeros(3), torch.zeros(3), torch.zeros(3)
    gs = [torch.tensor([1e-3, -2e-9, 5.0]), torch.tensor([2e-3, 1e-9, -1.0]), torch.tensor([0.0, 3e-9, 2.0])]
    pe, me, ve = np.zeros(3), np.zeros(3), np.zeros(3)
    for t, gt in enumerate(gs, 1):
        T.adam_tf_step(p, m, v, gt, t)
        gn = gt.numpy().astype(np.float64)
        me = 0.9 * me + 0.1 * gn
        ve = 0.999 * ve + 0.001 * gn * gn
        pe -= 1e-4 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * me / (np.sqrt(ve) + 1e-8)
    np.testing.assert_allclose(p.numpy(), pe, rtol=1e-4)


def test_confusion_matrix_and_iou():
    gt = np.array([[0, 0, 1, 1, 1]])
    pr = np.array([[0, 1, 1, 1, 0]])
    cm = T.confusion_matrix(gt, pr)
    assert cm.tolist() == [[1, 1], [1, 2]] and cm.dtype == np.int64
    assert T.iou_road(cm) == pytest.approx(2 / 4)


def test_variable_inventory():
    shapes = variable_shapes(3, 2)
    assert len(shapes) == 40   # 20 layers x (weights, biases); SURVEY says 38, miscounted
    assert sum(int(np.prod(s)) for s in shapes.values()) == 138_873_924
    assert list(shapes)[-1] == "conv_t3/bias"


@pytest.fixture(scope="module")
def small_model():
    variables = init_variables(cin=3, ncls=2, fc=128, seed=1234, init="ref")
    return variables


def test_initial_loss_is_ln2_and_logit_scale(small_model):
    x, lab = synthetic_batch(1, 32, 64, seed=0)
    o = FCN8sOracle(small_model)
    pred, logits = o.forward(x)
    assert pred.shape == (1, 32, 64, 1) and pred.dtype == torch.int64
    assert logits.shape == (1, 32, 64, 2)
    loss = float(o.loss(logits, lab))
    assert abs(loss - math.log(2.0)) < 1e-4


def test_fd_gradient_check_small_graph():
    """fp64 finite differences through conv->relu->pool->deconv->xent built from tf_ops."""
    torch.manual_seed(0)
    x = torch.randn(1, 4, 4, 2, dtype=torch.float64)
    w1 = torch.randn(3, 3, 2, 3, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(4, 4, 2, 3, dtype=torch.float64, requires_grad=True)
    lab = torch.nn.functional.one_hot(torch.randint(0, 2, (1, 4, 4)), 2).double()

    def f(w1_, wt_):
        a = T.relu(T.conv2d_same(x, w1_))
        p = T.max_pool_2x2(a)
        lg = T.conv2d_transpose_same(p, wt_, (4, 4), 2)
        return T.softmax_cross_entropy_with_logits(lg, lab).mean()

    assert torch.autograd.gradcheck(f, (w1, wt), eps=1e-6, atol=1e-5)


def test_paste_mask_restatement_matches_pil():
    """FCN.py:203-211 uses scipy.misc.toimage + PIL Image.paste(mask=mask): pin the NumPy restatement to PIL."""
    from PIL import Image
    rng = np.random.default_rng(3)
    for c in (3, 4):
        img = rng.integers(0, 256, (20, 33, c), dtype=np.uint8)
        prob = rng.random((20, 33)).astype(np.float32)
        seg = (prob > 0.5).reshape(20, 33, 1)
        mask = np.dot(seg, np.array([[0, 255, 0, 127]])).astype(np.uint8)          # FCN.py:207
        m = Image.fromarray(mask, mode="RGBA")
        im = Image.fromarray(img, mode="RGB" if c == 3 else "RGBA")
        im.paste(m, box=None, mask=m)                                              # FCN.py:209
        assert np.array_equal(T.paste_mask(img, prob), np.array(im))


def test_atrous_conv_equals_conv_with_zero_inserted_filter():
    """tf.nn.atrous_conv2d == conv2d with the filter up-sampled by inserting rate-1 zeros between taps (its definition)."""
    rng = np.random.default_rng(20)
    x = torch.tensor(rng.standard_normal((2, 9, 11, 3)).astype(np.float32))
    w = torch.tensor(rng.standard_normal((3, 3, 3, 4)).astype(np.float32))
    for rate in (1, 2, 3):
        k = 3 + 2 * (rate - 1)
        wz = torch.zeros((k, k, 3, 4))
        wz[::rate, ::rate] = w
        assert torch.allclose(T.atrous_conv2d_same(x, w, rate), T.conv2d_same(x, wz), atol=1e-5)


def test_resize_bilinear_align_corners_against_the_formula():
    rng = np.random.default_rng(21)
    x = rng.standard_normal((1, 4, 5, 2)).astype(np.float32)
    oh, ow = 7, 11
    got = T.resize_bilinear_align_corners(torch.tensor(x), (oh, ow)).numpy()
    ref = np.zeros((1, oh, ow, 2), np.float32)
    for oy in range(oh):
        sy = oy * (4 - 1) / (oh - 1); y0 = int(np.floor(sy)); y1 = min(y0 + 1, 3); fy = sy - y0
        for ox in range(ow):
            sx = ox * (5 - 1) / (ow - 1); x0 = int(np.floor(sx)); x1 = min(x0 + 1, 4); fx = sx - x0
            top = x[0, y0, x0] + (x[0, y0, x1] - x[0, y0, x0]) * fx
            bot = x[0, y1, x0] + (x[0, y1, x1] - x[0, y1, x0]) * fx
            ref[0, oy, ox] = top + (bot - top) * fy
    assert np.allclose(got, ref, atol=1e-5)
    assert np.allclose(T.global_avg_pool(torch.tensor(x)).numpy(), x.mean(axis=(1, 2)), atol=1e-6)
    assert np.allclose(T.avg_pool_2x2(torch.tensor(x[:, :4, :4])).numpy()[0, 0, 0], x[0, :2, :2].mean(axis=(0, 1)), atol=1e-6)


def test_windowed_pools_and_depthwise_conv_against_naive_loops():
    """The op-family oracles added for SURVEY §8f row 4 (avg / max pooling windows, depthwise conv) against direct loops, and
    the TF SAME geometry they share: out = ceil(in / s), pad_before = max((out - 1) s + k_eff - in, 0) // 2."""
    from oracle import tf_ops as T
    assert T._same_pad(5, 3, 2) == (3, 1, 1) and T._same_pad(6, 3, 2) == (3, 0, 1) and T._same_pad(7, 5, 1) == (7, 2, 2)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 7, 9, 3)).astype(np.float32)
    # average pool 3x2 windows, stride (2, 3), VALID
    y = T.avg_pool_valid(torch.tensor(x), 3, 2, 2, 3).numpy()
    assert y.shape == (2, 3, 3, 3)
    for oy in range(3):
        for ox in range(3):
            np.testing.assert_allclose(y[:, oy, ox], x[:, 2 * oy:2 * oy + 3, 3 * ox:3 * ox + 2].mean(axis=(1, 2)), rtol=1e-5, atol=1e-6)
    # max pool 3x3 / 2 SAME with ties: first maximum in row-major window order, padding never wins
    xi = rng.integers(0, 3, (1, 5, 6, 2)).astype(np.float32)
    ym, idx = T.max_pool_general(torch.tensor(xi), 3, 3, 2, "SAME")
    oh, pt, _ = T._same_pad(5, 3, 2)
    ow, pl, _ = T._same_pad(6, 3, 2)
    assert tuple(ym.shape) == (1, oh, ow, 2)
    dx = T.max_pool_general_grad(torch.ones_like(ym), idx, (5, 6), 3, 3, 2, "SAME").numpy()
    want = np.zeros_like(xi)
    for oy in range(oh):
        for ox in range(ow):
            for c in range(2):
                best, bk = None, None
                for ky in range(3):
                    for kx in range(3):
                        iy, ix = oy * 2 - pt + ky, ox * 2 - pl + kx
                        if 0 <= iy < 5 and 0 <= ix < 6 and (best is None or xi[0, iy, ix, c] > best):
                            best, bk = xi[0, iy, ix, c], (ky, kx, iy, ix)
                assert float(ym[0, oy, ox, c]) == best and int(idx[0, oy, ox, c]) == bk[0] * 3 + bk[1]
                want[0, bk[2], bk[3], c] += 1
    assert np.array_equal(dx, want)
    # depthwise 3x3, stride 2, and rate 2
    w = rng.standard_normal((3, 3, 3)).astype(np.float32)
    for s, r in ((2, 1), (1, 2)):
        y = T.depthwise_conv2d_same(torch.tensor(x), torch.tensor(w), s, r).numpy()
        oh, pt, _ = T._same_pad(7, 2 * r + 1, s)
        ow, pl, _ = T._same_pad(9, 2 * r + 1, s)
        ref = np.zeros((2, oh, ow, 3), np.float32)
        for oy in range(oh):
            for ox in range(ow):
                for ky in range(3):
                    for kx in range(3):
                        iy, ix = oy * s - pt + ky * r, ox * s - pl + kx * r
                        if 0 <= iy < 7 and 0 <= ix < 9:
                            ref[:, oy, ox] += x[:, iy, ix] * w[ky, kx]
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)
