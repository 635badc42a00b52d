"""Naive NumPy loop restatements of the TF ops (tiny shapes only).  TEST INFRASTRUCTURE ONLY.

These follow the op *definitions* in SURVEY.md Appendix B directly (explicit index
arithmetic, fp64), and exist to cross-check the torch mapping in oracle/tf_ops.py, which
has no upstream pin (PARITY UNPINNED).
"""
import numpy as np


def conv2d_same_naive(x, w, stride=1):
    """tf.nn.conv2d SAME (FCN.py:130): out=ceil(in/s), pad_before=pad_total//2."""
    n, h, wd, ci = x.shape
    kh, kw, _, co = w.shape
    oh, ow = -(-h // stride), -(-wd // stride)
    pt = max((oh - 1) * stride + kh - h, 0) // 2
    pl = max((ow - 1) * stride + kw - wd, 0) // 2
    y = np.zeros((n, oh, ow, co), np.float64)
    for b in range(n):
        for i in range(oh):
            for j in range(ow):
                for ky in range(kh):
                    for kx in range(kw):
                        yy, xx = i * stride - pt + ky, j * stride - pl + kx
                        if 0 <= yy < h and 0 <= xx < wd:
                            y[b, i, j] += x[b, yy, xx].astype(np.float64) @ w[ky, kx].astype(np.float64)
    return y


def conv2d_transpose_same_naive(x, w, out_hw, stride):
    """tf.nn.conv2d_transpose SAME (FCN.py:106,155): scatter form,
    y[n, i*s-p+ky, j*s-p+kx, co] += x[n,i,j,ci] * W[ky,kx,co,ci]."""
    n, h, wd, ci = x.shape
    kh, kw, co, _ = w.shape
    oh, ow = out_hw
    pt = max((-(-oh // stride) - 1) * stride + kh - oh, 0) // 2
    pl = max((-(-ow // stride) - 1) * stride + kw - ow, 0) // 2
    y = np.zeros((n, oh, ow, co), np.float64)
    for b in range(n):
        for i in range(h):
            for j in range(wd):
                for ky in range(kh):
                    for kx in range(kw):
                        yy, xx = i * stride - pt + ky, j * stride - pl + kx
                        if 0 <= yy < oh and 0 <= xx < ow:
                            y[b, yy, xx] += w[ky, kx].astype(np.float64) @ x[b, i, j].astype(np.float64)
    return y


def max_pool_2x2_naive(x):
    """2x2/s2 VALID max pool with first-max index (strict > scan in (dy,dx) row-major)."""
    n, h, w, c = x.shape
    oh, ow = h // 2, w // 2
    y = np.zeros((n, oh, ow, c), x.dtype)
    idx = np.zeros((n, oh, ow, c), np.uint8)
    for b in range(n):
        for i in range(oh):
            for j in range(ow):
                for ch in range(c):
                    best, bi = x[b, 2 * i, 2 * j, ch], 0
                    for k in range(1, 4):
                        v = x[b, 2 * i + k // 2, 2 * j + k % 2, ch]
                        if v > best:
                            best, bi = v, k
                    y[b, i, j, ch], idx[b, i, j, ch] = best, bi
    return y, idx


def softmax_xent_naive(logits, onehot):
    z = logits.astype(np.float64)
    z = z - z.max(-1, keepdims=True)
    lse = np.log(np.exp(z).sum(-1, keepdims=True))
    loss = (onehot * (lse - z)).sum(-1)
    grad = np.exp(z) / np.exp(z).sum(-1, keepdims=True) - onehot
    return loss, grad
