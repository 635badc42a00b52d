"""Hand-written known-answer vectors for the TF-1.x op semantics the oracle restates.

TensorFlow is not installable here and the reference ships no fixtures, so the oracle cannot be pinned against outputs
of the reference itself ("parity unpinned").  These cases pin it instead against the PUBLISHED definitions of
the ops -- each value below is worked out by hand from the formula in the TensorFlow API documentation cited in the
test, not produced by any code in this repository:

  * SAME padding:      tf.nn.convolution, "Padding" notes:  out = ceil(in / stride);
                       pad_total = max((out - 1) * stride + k - in, 0);  pad_before = pad_total // 2  (the odd one goes after)
  * conv2d_transpose:  tf.nn.conv2d_transpose: "the transpose (gradient) of conv2d"
  * max_pool:          tf.nn.max_pool VALID 2x2 / stride 2; ties route the gradient to the first element in window scan order
                       (maxpooling_op.cc SpatialMaxPoolWithArgMaxHelper keeps the first maximum: strict '>' update)
  * Adam:              tf.train.AdamOptimizer docs:  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1 - b1) g;
                       v = b2 v + (1 - b2) g^2;  var -= lr_t * m / (sqrt(v) + eps)
  * softmax xent:      tf.nn.softmax_cross_entropy_with_logits: -sum_c labels_c log softmax(logits)_c
  * reduce_max grad:   math_grad._MinOrMaxGrad: dy * indicators / num_selected
  * resize_bilinear:   align_corners=True: src = dst * (in - 1) / (out - 1)
  * dropout:           tf.nn.dropout: kept elements scaled by 1 / keep_prob
"""
import math

import numpy as np
import torch

from oracle import tf_ops as T


def test_same_padding_rule_including_the_asymmetric_cases():
    # (in, k, s) -> (out, before, after), each by hand from the documented rule
    cases = {(5, 3, 1): (5, 1, 1), (5, 2, 1): (5, 0, 1), (6, 4, 2): (3, 1, 1), (7, 4, 2): (4, 1, 2), (5, 7, 1): (5, 3, 3),
             (160, 16, 8): (20, 4, 4), (18, 7, 1): (18, 3, 3), (9, 3, 2): (5, 1, 1), (10, 3, 2): (5, 0, 1)}
    for (n, k, s), want in cases.items():
        assert T._same_pad(n, k, s) == want, ((n, k, s), T._same_pad(n, k, s), want)


def test_conv2d_same_known_answer():
    x = torch.arange(1, 10, dtype=torch.float32).reshape(1, 3, 3, 1)
    w = torch.ones(3, 3, 1, 1)
    want = [[12, 21, 16], [27, 45, 33], [24, 39, 28]]                 # 3x3 box sums with zero padding, by hand
    assert T.conv2d_same(x, w)[0, :, :, 0].tolist() == want
    # even kernel, stride 1: the extra padding column goes to the RIGHT (pad_before 0, pad_after 1)
    x1 = torch.tensor([1., 2., 3., 4., 5.]).reshape(1, 1, 5, 1)
    w1 = torch.ones(1, 2, 1, 1)
    assert T.conv2d_same(x1, w1)[0, 0, :, 0].tolist() == [3, 5, 7, 9, 5]
    # 4x4 stride 2 on 4 pixels (1 before, 1 after): windows [-1..2], [1..4]
    x2 = torch.tensor([1., 2., 3., 4.]).reshape(1, 1, 4, 1)
    w2 = torch.tensor([1., 10., 100., 1000.]).reshape(1, 4, 1, 1)
    assert T.conv2d_same(x2, w2, stride=2)[0, 0, :, 0].tolist() == [10 * 1 + 100 * 2 + 1000 * 3, 1 * 2 + 10 * 3 + 100 * 4]


def test_conv2d_transpose_is_the_gradient_of_conv2d_known_answer():
    # k = 4, s = 2, SAME, one input pixel -> 2x2 output: y[oy][ox] = x * W[oy + 1][ox + 1]
    w = torch.arange(16, dtype=torch.float32).reshape(4, 4, 1, 1)      # [ky,kx,Cout,Cin]
    y = T.conv2d_transpose_same(torch.tensor([[[[2.0]]]]), w, (2, 2), 2)
    assert y[0, :, :, 0].tolist() == [[10, 12], [18, 20]]
    # two input pixels in a row: out[o] = sum_i x[i] * w[o - 2 i + 1]
    w1 = torch.tensor([1., 10., 100., 1000.]).reshape(1, 4, 1, 1)
    x1 = torch.tensor([1., 2.]).reshape(1, 1, 2, 1)
    y1 = T.conv2d_transpose_same(x1, w1, (1, 4), 2)
    assert y1[0, 0, :, 0].tolist() == [10, 100 + 2 * 1, 1000 + 2 * 10, 2 * 100]
    # and it is the adjoint of the strided conv: <conv(a), b> == <a, conv_transpose(b)>
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn((2, 8, 12, 3), generator=g), torch.randn((2, 4, 6, 5), generator=g)
    wt = torch.randn((4, 4, 3, 5), generator=g)                        # HWIO of the conv == [k,k,Cout,Cin] of its transpose
    lhs = (T.conv2d_same(a, wt, stride=2) * b).sum()
    rhs = (a * T.conv2d_transpose_same(b, wt, (8, 12), 2)).sum()
    assert abs(float(lhs - rhs)) <= 1e-3 * abs(float(lhs))


def test_max_pool_values_and_first_max_routing():
    x = np.array([[1, 5, 2, 2], [5, 3, 2, 2], [0, 0, 7, 1], [0, 0, 1, 7]], np.float32).reshape(1, 4, 4, 1)
    y, idx = T.max_pool_2x2_with_argmax(x)
    assert y[0, :, :, 0].tolist() == [[5, 2], [0, 7]]
    # window scan order (dy,dx): (0,0)=0 (0,1)=1 (1,0)=2 (1,1)=3; ties -> the first
    assert idx[0, :, :, 0].tolist() == [[1, 0], [0, 0]]
    dx = T.max_pool_2x2_grad(np.ones((1, 2, 2, 1), np.float32), idx, (4, 4))
    assert dx[0, :, :, 0].tolist() == [[0, 1, 1, 0], [0, 0, 0, 0], [1, 0, 1, 0], [0, 0, 0, 0]]


def test_adam_first_steps_known_answer():
    # from zero slots: m1 = 0.1 g, v1 = 0.001 g^2, lr_1 = lr sqrt(0.001) / 0.1 -> update = lr * g / (|g| + eps sqrt(1000))
    lr, eps = 1e-4, 1e-8
    g = torch.tensor([0.5, -2.0, 1e-3])
    p, m, v = torch.zeros(3), torch.zeros(3), torch.zeros(3)
    lr_t = T.adam_tf_step(p, m, v, g, t=1, lr=lr, eps=eps)
    assert abs(lr_t - lr * math.sqrt(0.001) / 0.1) <= 1e-9
    np.testing.assert_allclose(m.numpy(), 0.1 * g.numpy(), rtol=1e-6)
    # (1 - beta2 is formed in float32, as TF's ApplyAdam kernel does in the variable dtype: 0.0010000467 instead of 0.001)
    np.testing.assert_allclose(v.numpy(), 0.001 * g.numpy() ** 2, rtol=1e-4)
    want = -lr * g.numpy() / (np.abs(g.numpy()) + eps * math.sqrt(1000.0))
    np.testing.assert_allclose(p.numpy(), want, rtol=1e-4)
    # epsilon sits OUTSIDE the bias correction (TF), not inside (the paper's epsilon-hat): a tiny gradient shows the difference
    g2 = torch.tensor([1e-9])
    p2, m2, v2 = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    T.adam_tf_step(p2, m2, v2, g2, t=1, lr=lr, eps=eps)
    tf_form = -lr * 1e-9 / (1e-9 + eps * math.sqrt(1000.0))
    paper_form = -lr * 1e-9 / (1e-9 + eps)
    assert abs(float(p2) - tf_form) <= 1e-3 * abs(tf_form) and abs(float(p2) - paper_form) > 0.5 * abs(paper_form)


def test_softmax_cross_entropy_known_answers():
    z = torch.tensor([[0.0, 0.0], [0.0, math.log(3.0)], [math.log(3.0), 0.0]])
    lab = torch.tensor([[1.0, 0.0], [1.0, 0.0], [1.0, 0.0]])
    got = T.softmax_cross_entropy_with_logits(z, lab).numpy()
    np.testing.assert_allclose(got, [math.log(2.0), math.log(4.0), math.log(4.0 / 3.0)], rtol=1e-6)
    assert T.argmax_last(torch.tensor([[1.0, 1.0], [0.0, 2.0]])).tolist() == [0, 1]          # first index on ties


def test_reduce_max_gradient_shares_ties_and_bilinear_align_corners_and_dropout():
    x = torch.tensor([[1.0, 3.0], [3.0, 2.0]]).reshape(1, 2, 2, 1).requires_grad_()
    T.global_max_pool(x).backward(torch.tensor([[4.0]]))
    assert x.grad[0, :, :, 0].tolist() == [[0, 2], [2, 0]]                 # dy / num_selected on both maxima
    r = T.resize_bilinear_align_corners(torch.tensor([0.0, 1.0]).reshape(1, 1, 2, 1), (1, 3))
    assert r[0, 0, :, 0].tolist() == [0.0, 0.5, 1.0]
    r2 = T.resize_bilinear_align_corners(torch.tensor([0.0, 3.0, 6.0]).reshape(1, 1, 3, 1), (1, 5))
    np.testing.assert_allclose(r2[0, 0, :, 0].numpy(), [0.0, 1.5, 3.0, 4.5, 6.0], rtol=1e-6)
    d = T.dropout(torch.tensor([1.0, 2.0, 3.0, 4.0]), 0.8, torch.tensor([1.0, 0.0, 1.0, 1.0]))
    np.testing.assert_allclose(d.numpy(), [1.25, 0.0, 3.75, 5.0], rtol=1e-6)


def test_depthwise_conv_avg_pool_and_same_max_pool_known_answers():
    """tf.nn.depthwise_conv2d (documented: output[b,i,j,k] = sum_{di,dj} filter[di,dj,k,0] * input[b, s i + r di, s j + r dj, k],
    SAME padding by the padding rule above with the EFFECTIVE filter size (k - 1) r + 1), tf.nn.avg_pool VALID (windows that
    do not fit are dropped) and tf.nn.max_pool SAME (padding never wins), each by hand."""
    # depthwise 1x3 filter [1, 10, 100] per channel, two channels scaled 1 and 2, stride 1, rate 1: y[j] = x[j-1] + 10 x[j] + 100 x[j+1]
    x = torch.tensor([1., 2., 3., 4.]).reshape(1, 1, 4, 1).repeat(1, 1, 1, 2)
    w = torch.tensor([1., 10., 100.]).reshape(1, 3, 1) * torch.tensor([1., 2.])
    y = T.depthwise_conv2d_same(x, w, 1, 1)
    assert y[0, 0, :, 0].tolist() == [210., 321., 432., 43.] and y[0, 0, :, 1].tolist() == [420., 642., 864., 86.]
    # rate 2: taps at j-2, j, j+2 (effective size 5, SAME pads 2 + 2)
    y = T.depthwise_conv2d_same(x[..., :1], w[..., :1], 1, 2)
    assert y[0, 0, :, 0].tolist() == [310., 420., 31., 42.]
    # stride 2, in = 4, k = 3: out = 2, pad_total = (2 - 1) 2 + 3 - 4 = 1 -> before 0, after 1: windows start at 0 and 2
    y = T.depthwise_conv2d_same(x[..., :1], w[..., :1], 2, 1)
    assert y[0, 0, :, 0].tolist() == [321., 43.]                      # x0 + 10 x1 + 100 x2, x2 + 10 x3 + 100 * pad
    # avg_pool 2x2 / stride 2 VALID on 3x5: the last row / column do not fit and are dropped
    a = torch.arange(15, dtype=torch.float32).reshape(1, 3, 5, 1)
    assert T.avg_pool_valid(a, 2, 2, 2, 2)[0, :, :, 0].tolist() == [[3.0, 5.0]]
    # PSPNet's 1x3 window with stride (1, 3) on one 1x6 row (PSPNet.py:165)
    assert T.avg_pool_valid(a[:, :1, :, :].repeat(1, 1, 1, 1)[:, :, :3, :], 1, 3, 1, 3)[0, 0, :, 0].tolist() == [1.0]
    # max_pool 3x3 / stride 2 SAME on all-negative values: the zero-valued padding of a naive implementation would win; TF's does not
    m = -torch.arange(1, 17, dtype=torch.float32).reshape(1, 4, 4, 1)
    ym, idx = T.max_pool_general(m, 3, 3, 2, "SAME")                    # in 4, k 3, s 2: out 2, pad_total 1 -> before 0, after 1
    assert ym[0, :, :, 0].tolist() == [[-1., -3.], [-9., -11.]]
    assert idx[0, :, :, 0].tolist() == [[0, 0], [0, 0]]                 # the window's first element is its maximum here
