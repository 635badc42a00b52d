"""TF-1.15 op semantics restated with torch CPU fp32 ops.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference call site it follows (paths relative to the upstream
repo root).  Layouts are the reference's: activations NHWC, conv filters HWIO
(`[kh,kw,Cin,Cout]`), transposed-conv filters `[kh,kw,Cout,Cin]`.

PARITY UNPINNED (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _same_pad(in_size: int, k: int, s: int):
    """TF 'SAME' padding split: out=ceil(in/s); extra pad goes to bottom/right."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def conv2d_same(x: torch.Tensor, w: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """tf.nn.conv2d(x, W, strides=[1,s,s,1], padding='SAME')  (FCN.py:130).

    x: [N,H,W,Cin]; w: [kh,kw,Cin,Cout] -> [N,ceil(H/s),ceil(W/s),Cout]. Cross-correlation.
    """
    kh, kw = w.shape[0], w.shape[1]
    _, pt, pb = _same_pad(x.shape[1], kh, stride)
    _, pl, pr = _same_pad(x.shape[2], kw, stride)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_transpose_same(x: torch.Tensor, w: torch.Tensor, out_hw, stride: int) -> torch.Tensor:
    """tf.nn.conv2d_transpose(x, W[kh,kw,Cout,Cin], output_shape, [1,s,s,1], 'SAME')
    (FCN.py:106,155).  Defined as the input-gradient of the SAME conv2d whose input has
    `output_shape`:  y[n, i*s - p + ky, j*s - p + kx, co] += x[n,i,j,ci] * W[ky,kx,co,ci].
    """
    kh, kw = w.shape[0], w.shape[1]
    oh, ow = out_hw
    _, pt, _ = _same_pad(oh, kh, stride)
    _, pl, _ = _same_pad(ow, kw, stride)
    xn = x.permute(0, 3, 1, 2)
    # torch conv_transpose2d weight: [Cin, Cout, kh, kw]
    wt = w.permute(3, 2, 0, 1)
    full = F.conv_transpose2d(xn, wt, stride=stride)  # no padding: size (in-1)*s + k
    y = full[:, :, pt:pt + oh, pl:pl + ow]
    # when (in-1)*s+k-pt < oh the tail would be short; the hot-path layers (k=2s) never hit it
    assert y.shape[2] == oh and y.shape[3] == ow, (y.shape, oh, ow)
    return y.permute(0, 2, 3, 1).contiguous()


def atrous_conv2d_same(x: torch.Tensor, w: torch.Tensor, rate: int) -> torch.Tensor:
    """tf.nn.atrous_conv2d(x, W[kh,kw,Cin,Cout], rate, padding='SAME')  (utils.py:210-231): the filter taps sit `rate`
    pixels apart; effective size k + (k-1)(rate-1), odd for odd k, so SAME pads rate*(k//2) on every side."""
    kh, kw = w.shape[0], w.shape[1]
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), dilation=rate, padding=(rate * (kh // 2), rate * (kw // 2)))
    return y.permute(0, 2, 3, 1).contiguous()


def resize_bilinear_align_corners(x: torch.Tensor, size) -> torch.Tensor:
    """tf.image.resize_bilinear(x, size, align_corners=True)  (utils.py:329-330): src = dst * (in-1)/(out-1)."""
    y = F.interpolate(x.permute(0, 3, 1, 2), size=tuple(size), mode="bilinear", align_corners=True)
    return y.permute(0, 2, 3, 1).contiguous()


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """tflearn global_avg_pool (utils.py:312-313): reduce_mean over H and W -> [N, C]."""
    return x.mean(dim=(1, 2))


def global_max_pool(x: torch.Tensor) -> torch.Tensor:
    """tflearn global_max_pool (utils.py:315-316): reduce_max over H and W -> [N, C].  Its TF gradient divides dy
    equally among the maximal elements (math_grad._MinOrMaxGrad: indicators / num_selected); torch.amax does the same."""
    return x.amax(dim=(1, 2))


def zero_padding(x: torch.Tensor, pad: int) -> torch.Tensor:
    """Zero_Padding (utils.py:325-327): tf.pad with `pad` zeros on every side of H and W."""
    return F.pad(x, (0, 0, pad, pad, pad, pad))


def avg_pool_2x2(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.avg_pool(ksize 2x2, stride 2, 'VALID')  (utils.py:309)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def bias_add(x: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tf.nn.bias_add (FCN.py:107,132,157)."""
    return x + b


def relu(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.relu (FCN.py:134)."""
    return torch.relu(x)


def max_pool_2x2(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.max_pool(ksize 2x2, stride 2, 'VALID') (FCN.py:163). Autograd routes the
    gradient to the first maximal element in row-major window order, as TF's MaxPoolGrad."""
    y = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2)
    return y.permute(0, 2, 3, 1).contiguous()


def max_pool_2x2_with_argmax(x: np.ndarray):
    """NumPy restatement of the 2x2/s2 VALID max-pool with the in-window index (0..3,
    row-major (dy,dx)) of the FIRST maximal element (strict '>' scan), i.e. the element
    TF's MaxPoolGrad routes the gradient to."""
    n, h, w, c = x.shape
    oh, ow = h // 2, w // 2
    win = np.stack([x[:, 0:2 * oh:2, 0:2 * ow:2], x[:, 0:2 * oh:2, 1:2 * ow:2],
                    x[:, 1:2 * oh:2, 0:2 * ow:2], x[:, 1:2 * oh:2, 1:2 * ow:2]], axis=0)
    best = win[0].copy()
    idx = np.zeros(best.shape, np.uint8)
    for k in range(1, 4):
        gt = win[k] > best
        best = np.where(gt, win[k], best)
        idx = np.where(gt, np.uint8(k), idx)
    return best, idx


def max_pool_2x2_grad(dy: np.ndarray, idx: np.ndarray, in_hw) -> np.ndarray:
    """MaxPoolGrad from the stored in-window index."""
    n, oh, ow, c = dy.shape
    h, w = in_hw
    dx = np.zeros((n, h, w, c), dy.dtype)
    for k in range(4):
        ky, kx = divmod(k, 2)
        dx[:, ky:2 * oh:2, kx:2 * ow:2] = np.where(idx == k, dy, 0)
    return dx


def dropout(x: torch.Tensor, keep_prob: float, mask: torch.Tensor | None) -> torch.Tensor:
    """tf.nn.dropout(x, keep_prob) = x * floor(keep_prob + U[0,1)) / keep_prob (FCN.py:167).
    TF's RNG stream is not reproducible, so the keep-mask (0/1) is injected."""
    if keep_prob >= 1.0 or mask is None:
        return x
    return x * mask / keep_prob


def softmax_cross_entropy_with_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """tf.nn.softmax_cross_entropy_with_logits(logits=, labels=) over the last axis
    (FCN.py:334): loss = sum_c labels_c * (log sum exp z - z_c), z = logits - max."""
    z = logits - logits.max(dim=-1, keepdim=True).values
    lse = torch.log(torch.exp(z).sum(dim=-1, keepdim=True))
    return (labels * (lse - z)).sum(dim=-1)


def argmax_last(logits: torch.Tensor) -> torch.Tensor:
    """tf.argmax(logits, 3): int64, first index on ties (FCN.py:111)."""
    return torch.argmax(logits, dim=-1)


def confusion_matrix(gt: np.ndarray, pred: np.ndarray, ncls: int = 2) -> np.ndarray:
    """cm[gt,pred] int64 (new functionality)."""
    return np.bincount((gt.astype(np.int64) * ncls + pred.astype(np.int64)).ravel(),
                       minlength=ncls * ncls).reshape(ncls, ncls).astype(np.int64)


def iou_road(cm: np.ndarray) -> float:
    d = cm[1, 1] + cm[0, 1] + cm[1, 0]
    return float(cm[1, 1]) / float(d) if d else 0.0


def paste_mask(image_u8: np.ndarray, road_prob: np.ndarray, color=(0, 255, 0, 127)) -> np.ndarray:
    """paste_mask (FCN.py:203-211) restated with NumPy: segmentation = prob > 0.5; the RGBA colour is
    pasted with itself as mask, i.e. PIL's integer alpha blend t = dst*(255-a) + src*a,
    out = ((t+128) + ((t+128) >> 8)) >> 8 on every channel of the image (checked against PIL itself in
    tests/test_oracle.py)."""
    seg = road_prob > 0.5
    img = image_u8.astype(np.int64)
    a = int(color[3])
    out = image_u8.copy()
    for c in range(image_u8.shape[-1]):
        t = img[..., c] * (255 - a) + int(color[c]) * a + 128
        out[..., c] = np.where(seg, (t + (t >> 8)) >> 8, img[..., c]).astype(np.uint8)
    return out


def adam_tf_step(p, m, v, g, t: int, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer ApplyAdam (FCN.py:338-340), TF formula (epsilon outside the
    bias correction): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
    Operates in-place on fp32 torch tensors; t starts at 1."""
    lr_t = np.float32(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    # TF's ApplyAdam kernel evaluates the same recurrences in this form, in the variable dtype:
    #   m += (g - m) * (1 - b1);  v += (g*g - v) * (1 - b2);  var -= (m * lr_t) / (sqrt(v) + eps)
    c1 = float(np.float32(1.0) - np.float32(b1))
    c2 = float(np.float32(1.0) - np.float32(b2))
    m.add_((g - m) * c1)
    v.add_((g * g - v) * c2)
    p.sub_(lr_t * m / (v.sqrt() + eps))
    return float(lr_t)


def momentum_tf_step(p, a, g, lr, mu):
    """tf.train.MomentumOptimizer: a = mu*a + g; p -= lr*a."""
    a.mul_(mu).add_(g)
    p.sub_(lr * a)


def to_bf16_grid(x: torch.Tensor) -> torch.Tensor:
    """Round an fp32 tensor to the nearest bf16 value (kept as fp32).  Used to build the
    'bf16-storage' variant of the oracle that mirrors where the CUDA path rounds."""
    return x.to(torch.bfloat16).to(torch.float32)


# ---- remaining op families -----------------------------------------------------------------------
def avg_pool_valid(x: torch.Tensor, kh: int, kw: int, sh: int, sw: int) -> torch.Tensor:
    """tf.nn.avg_pool(x, [1,kh,kw,1], [1,sh,sw,1], 'VALID')  (utils.py:309; PSPNet.py:147-165 pyramid windows)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), (kh, kw), (sh, sw)).permute(0, 2, 3, 1).contiguous()


def max_pool_general(x: torch.Tensor, kh: int, kw: int, stride: int, padding: str = "VALID"):
    """tf.nn.max_pool(x, [1,kh,kw,1], [1,s,s,1], padding)  (utils.py:306; PSPNet.py:34,190).  Returns (y, idx) with idx the
    position ky*kw + kx of the FIRST maximum of every window (the element TF's MaxPoolGrad routes the gradient to); SAME
    padding never wins the max."""
    n, h, w, c = x.shape
    if padding == "SAME":
        oh, pt, pb = _same_pad(h, kh, stride)
        ow, pl, pr = _same_pad(w, kw, stride)
    else:
        oh, ow, pt, pb, pl, pr = (h - kh) // stride + 1, (w - kw) // stride + 1, 0, 0, 0, 0
    xp = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float("-inf"))
    win = xp.unfold(2, kh, stride).unfold(3, kw, stride)[:, :, :oh, :ow]            # [n, c, oh, ow, kh, kw]
    flat = win.reshape(n, c, oh, ow, kh * kw)
    y = flat.max(dim=4).values
    first = (flat == y.unsqueeze(4)).to(torch.uint8).argmax(dim=4)                   # first index attaining the max
    return y.permute(0, 2, 3, 1).contiguous(), first.permute(0, 2, 3, 1).to(torch.uint8).contiguous()


def max_pool_general_grad(dy: torch.Tensor, idx: torch.Tensor, in_hw, kh: int, kw: int, stride: int, padding: str = "VALID"):
    """MaxPoolGrad from the first-max positions: dx[n, oy*s - pt + ky, ox*s - pl + kx, c] += dy[n, oy, ox, c]."""
    n, oh, ow, c = dy.shape
    h, w = in_hw
    pt = _same_pad(h, kh, stride)[1] if padding == "SAME" else 0
    pl = _same_pad(w, kw, stride)[1] if padding == "SAME" else 0
    dx = torch.zeros((n, h, w, c), dtype=dy.dtype)
    ky, kx = (idx // kw).long(), (idx % kw).long()
    oy = torch.arange(oh).view(1, oh, 1, 1)
    ox = torch.arange(ow).view(1, 1, ow, 1)
    iy, ix = oy * stride - pt + ky, ox * stride - pl + kx
    nn_ = torch.arange(n).view(n, 1, 1, 1).expand_as(iy)
    cc = torch.arange(c).view(1, 1, 1, c).expand_as(iy)
    dx.index_put_((nn_.reshape(-1), iy.reshape(-1), ix.reshape(-1), cc.reshape(-1)), dy.reshape(-1), accumulate=True)
    return dx


def depthwise_conv2d_same(x: torch.Tensor, w: torch.Tensor, stride: int = 1, rate: int = 1) -> torch.Tensor:
    """tf.nn.depthwise_conv2d(x, filter[kh,kw,C,1], [1,s,s,1], 'SAME', rate=[r,r])  (DeepLabv3Plus.py:49,
    EfficientNet.py:173,453): w [kh,kw,C]; effective window (k-1) r + 1."""
    kh, kw, c = w.shape
    _, pt, pb = _same_pad(x.shape[1], (kh - 1) * rate + 1, stride)
    _, pl, pr = _same_pad(x.shape[2], (kw - 1) * rate + 1, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=rate, groups=c)
    return y.permute(0, 2, 3, 1).contiguous()


def swish(x: torch.Tensor) -> torch.Tensor:
    """x * sigmoid(x)  (EfficientNet.py's activation)."""
    return x * torch.sigmoid(x)


def channel_scale(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """x [N,H,W,C] * s [N,C] broadcast over H, W (the squeeze-excite multiply of EfficientNet.py's SE block)."""
    return x * s[:, None, None, :]
