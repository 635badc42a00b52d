"""Per-op Python wrappers over the C ABI (`include/segk.h`).  Tensors are torch CUDA tensors
used purely as device buffers; every call enqueues on torch's current stream.

Layouts: activations NHWC bf16, logits NHWC fp32, images NHWC u8, master weights fp32 in
the reference's HWIO / [k,k,Cout,Cin] order."""
from __future__ import annotations

import torch

from ._lib import Context, EPI_OUT_F32, EPI_RELU, DT_BF16, DT_U8, DT_F32

_CTX = {}


def context(device=None) -> Context:
    if device is None:
        device = torch.cuda.current_device()
    if isinstance(device, torch.device):
        device = device.index or 0
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def _p(t):
    return 0 if t is None else t.data_ptr()


def _pitch(t):
    """Channel pitch of an NHWC tensor that is a channel slice of a wider one (zero-copy Concat view, segk_set_pitch);
    0 for a dense tensor."""
    if t is None or t.is_contiguous():
        return 0
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    if sc != 1 or sw < c or sh != w * sw or sn != h * sh or sw % 8 or (t.data_ptr() & 15):
        raise ValueError(f"not a channel-slice view: shape {tuple(t.shape)} strides {tuple(t.stride())}")
    return sw


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dt(x):
    if x.dtype == torch.bfloat16:
        return DT_BF16
    if x.dtype == torch.uint8:
        return DT_U8
    if x.dtype == torch.float32:
        return DT_F32
    raise TypeError(f"unsupported dtype {x.dtype}")


def _valid_frac(n, k):
    """Average fraction of a k-tap SAME window that lies inside [0, n) (1-D)."""
    p = k // 2
    return sum(min(o + k - 1 - p, n - 1) - max(o - p, 0) + 1 for o in range(n)) / float(n * k)


_VF = {}


def conv_flops(n, h, w, cin, cout, kh, kw):
    """Algorithmic (valid-tap) FLOPs of one conv pass: multiplies against SAME padding excluded."""
    key = (h, w, kh, kw)
    if key not in _VF:
        _VF[key] = _valid_frac(h, kh) * _valid_frac(w, kw)
    return 2.0 * n * h * w * kh * kw * cin * cout * _VF[key]


def deconv_flops(n, h, w, cin, cout, k, s):
    vf = (s * h) / float((h - 1) * s + k) * (s * w) / float((w - 1) * s + k)
    return 2.0 * n * h * w * k * k * cin * cout * vf


class Profile:
    """Per-launch CUDA-event timing of the C-ABI calls, recorded on the launching stream."""

    def __init__(self):
        self.records = []     # (family, work, unit, start_event, end_event)

    def summary(self, detail=False):
        """Aggregate by C entry point (or by entry point + integer arguments when detail)."""
        out = {}
        for fam, work, unit, e0, e1, dims in self.records:
            key = fam + str(list(dims)) if detail else fam
            d = out.setdefault(key, {"launches": 0, "ms": 0.0, "work": 0.0, "unit": unit})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["work"] += work
        return out


class Ops:
    """Bound to one device context."""

    def __init__(self, device=None):
        self.ctx = context(device)
        self._call = self.ctx.call
        self.profile = None
        self._work = None

    def call(self, name, *args):
        if self.profile is None:
            return self._call(name, *args)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self._call(name, *args)
        e1.record()
        work, unit = self._work if self._work is not None else (0.0, "")
        self._work = None
        dims = tuple(a for a in args[:-1] if isinstance(a, int) and not isinstance(a, bool) and 0 <= a < (1 << 20))
        self.profile.records.append((name, work, unit, e0, e1, dims))

    def _views(self, tin=None, tout=None):
        """Announce channel-slice views (zero-copy Concat) to the next C call: `tin` its gradient / input operand, `tout`
        its output."""
        pi, po = _pitch(tin), _pitch(tout)
        if pi or po:
            self._call("segk_set_pitch", pi, po)

    def _w(self, work, unit):
        if self.profile is not None:
            self._work = (work, unit)

    # ---- weight packing: kernel layouts are blocked by 64-wide k-chunks, [T][ceil(K/64)][rows][64] ------
    def pack_conv_weights(self, w, wk=None, wd=None):
        kh, kw, cin, cout = w.shape
        dev = w.device
        if wk is None:
            wk = torch.empty((kh * kw, -(-cin // 64), cout, 64), dtype=torch.bfloat16, device=dev)
        if wd is None:
            wd = torch.empty((kh * kw, -(-cout // 64), cin, 64), dtype=torch.bfloat16, device=dev)
        self._w(8.0 * w.numel(), "byte")
        self.call("segk_pack_conv_weights", _p(w), _p(wk), _p(wd), kh, kw, cin, cout, _stream())
        return wk, wd

    def pack_deconv_weights(self, w, s, wk=None, wd=None):
        k, _, cout, cin = w.shape
        dev = w.device
        if wk is None:
            wk = torch.empty((s * s * 4, -(-cin // 64), cout, 64), dtype=torch.bfloat16, device=dev)
        if wd is None:
            wd = torch.empty((k * k, -(-cout // 64), cin, 64), dtype=torch.bfloat16, device=dev)
        self._w(8.0 * w.numel(), "byte")
        self.call("segk_pack_deconv_weights", _p(w), _p(wk), _p(wd), k, s, cin, cout, _stream())
        return wk, wd

    def pack_matrix(self, w3, cp=None, tr=None):
        """w3 fp32 [T][A][B] -> cp bf16 [T][ceil(B/64)][A][64] (rows A, k = B), tr bf16 [T][ceil(A/64)][B][64]."""
        t, a, b = w3.shape
        if cp is None:
            cp = torch.empty((t, -(-b // 64), a, 64), dtype=torch.bfloat16, device=w3.device)
        if tr is None:
            tr = torch.empty((t, -(-a // 64), b, 64), dtype=torch.bfloat16, device=w3.device)
        self._w(8.0 * w3.numel(), "byte")
        self.call("segk_pack_matrix", _p(w3), _p(cp), _p(tr), t, a, b, _stream())
        return cp, tr

    # ---- phase-packed transposed conv with tiny Cout (conv_t3): one 4-tap GEMM over the block grid ----
    def pack_deconv_packed(self, w, s, bf=None, bt=None):
        k, _, cout, cin = w.shape
        r = s * s * cout
        if bf is None:
            bf = torch.empty((4, cin // 64, r, 64), dtype=torch.bfloat16, device=w.device)
        if bt is None:
            bt = torch.empty((4, r // 64, cin, 64), dtype=torch.bfloat16, device=w.device)
        self.call("segk_pack_deconv_packed", _p(w), _p(bf), _p(bt), k, s, cin, cout, _stream())
        return bf, bt

    def deconv2d_packed_fwd(self, x, bf, bias, y, k, s):
        n, h, w, cin = x.shape
        cout = y.shape[3]
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_packed_fwd", _p(x), _p(bf), _p(bias), _p(y), n, h, w, cin, cout, k, s, _stream())
        return y

    def deconv_pack_dy(self, dy, dyb, s):
        n, hb, wb, r = dyb.shape
        cout = dy.shape[3]
        self._w(float(dy.numel() * dy.element_size() + 2 * dyb.numel()), "byte")
        self.call("segk_deconv_pack_dy", _p(dy), int(dy.dtype == torch.float32), _p(dyb), n, hb - 1, wb - 1, cout, s, _stream())
        return dyb

    def deconv2d_packed_dgrad(self, dyb, bt, dx, cout, k, s, relu_mask=None, colsum=None):
        n, h, w, cin = dx.shape
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_packed_dgrad", _p(dyb), _p(bt), _p(relu_mask), _p(dx), _p(colsum), n, h, w, cin, cout, k, s,
                  _stream())
        return dx

    def deconv2d_packed_wgrad(self, x, dyb, dw, dwt, k, s, accumulate=False):
        """dw [k,k,Cout,Cin] fp32; dwt: fp32 scratch [4, Cin, s*s*Cout]."""
        n, h, w, cin = x.shape
        cout = dw.shape[2]
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_packed_wgrad", _p(x), _p(dyb), _p(dwt), n, h, w, cin, cout, k, s, _stream())
        self.call("segk_deconv_unpack_dw", _p(dwt), _p(dw), k, s, cin, cout, int(accumulate), _stream())
        return dw

    def pack_im2col_weights(self, w, wk=None):
        kh, kw, cin, cout = w.shape
        if wk is None:
            wk = torch.empty((1, cout, 64), dtype=torch.bfloat16, device=w.device)
        self.call("segk_pack_im2col_weights", _p(w), _p(wk), kh * kw * cin, cout, _stream())
        return wk

    # ---- first layer: patches built in shared memory, no patch tensor (firstconv.cu) ------------
    def conv2d_first_fwd(self, x, wk, bias, y, kh, kw, relu=True, relu_bits=None):
        n, h, w, cin = x.shape
        cout = y.shape[3]
        self._w(float(x.numel() * x.element_size() + 2 * y.numel()), "byte")
        self.call("segk_conv2d_first_fwd", _p(x), _dt(x), _p(wk), _p(bias), _p(y), _p(relu_bits), n, h, w, cin, cout, kh, kw,
                  EPI_RELU if relu else 0, _stream())
        return y

    def conv2d_first_wgrad(self, x, dy, dw, kh, kw, dbias=None):
        n, h, w, cin = x.shape
        cout = dy.shape[3]
        self._w(float(x.numel() * x.element_size() + 2 * dy.numel()), "byte")
        self.call("segk_conv2d_first_wgrad", _p(x), _dt(x), _p(dy), _p(dw), _p(dbias), n, h, w, cin, cout, kh, kw,
                  _stream())
        return dw

    # ---- patch-space helpers (conv1_1 / conv_t3 as tensor-core GEMMs) --------------------------
    def im2col_k64(self, x, P, kh, kw):
        n, h, w, cin = x.shape
        self._w(float(x.numel() * x.element_size() + 2 * P.numel()), "byte")
        self.call("segk_im2col_k64", _p(x), _dt(x), _p(P), n, h, w, cin, kh, kw, _stream())
        return P

    def deconv_patch_gather(self, dy, P, k, s):
        n, h, w, _ = P.shape
        cout = dy.shape[3]
        self._w(float(dy.numel() * dy.element_size() + 2 * P.numel()), "byte")
        self.call("segk_deconv_patch_gather", _p(dy), int(dy.dtype == torch.float32), _p(P), n, h, w, cout, k, s,
                  _stream())
        return P

    def deconv_col2im(self, yp, bias, y, k, s, residual=None):
        n, h, w, _ = yp.shape
        cout = y.shape[3]
        self._w(float(4 * yp.numel() + y.numel() * y.element_size()), "byte")
        self.call("segk_deconv_col2im", _p(yp), _p(bias), _p(residual), _p(y), int(y.dtype == torch.float32), n, h, w,
                  cout, k, s, _stream())
        return y

    # ---- tensor-core conv family ------------------------------------------------------
    def conv2d_fwd(self, x, wk, bias, y, kh, kw, relu=True, residual=None, flops=None, relu_bits=None):
        """`relu_bits` (int32 [N,H,W,Cout/32]): 1-bit ReLU mask of y written by the same epilogue."""
        n, h, w, cin = x.shape
        cout = y.shape[3]
        flags = (EPI_RELU if relu else 0) | (EPI_OUT_F32 if y.dtype == torch.float32 else 0)
        self._w(conv_flops(n, h, w, cin, cout, kh, kw) if flops is None else flops, "flop")
        self._views(tin=x, tout=y)
        self.call("segk_conv2d_fwd", _p(x), _p(wk), _p(bias), _p(residual), _p(y), _p(relu_bits), n, h, w, cin, cout, kh, kw,
                  flags, _stream())
        return y

    def conv2d_fwd_narrow(self, x, wk, bias, y, kh, kw, cout_padded, relu=False):
        """Conv to a few output channels on the tensor cores: wk / bias zero-padded to cout_padded channels, y fp32
        [N,H,W,out_cols] receives the first out_cols columns."""
        n, h, w, cin = x.shape
        self._w(conv_flops(n, h, w, cin, y.shape[3], kh, kw), "flop")
        self.call("segk_conv2d_fwd_narrow", _p(x), _p(wk), _p(bias), _p(y), y.shape[3], n, h, w, cin, int(cout_padded), kh, kw,
                  EPI_RELU if relu else 0, _stream())
        return y

    def pad_channels(self, src, dst):
        """dst bf16 [..., C] = src (fp32 / bf16) [..., c] zero-padded."""
        c, C = src.shape[-1], dst.shape[-1]
        self.call("segk_pad_channels", _p(src), int(src.dtype == torch.float32), _p(dst), src.numel() // c, c, C, _stream())
        return dst

    def relu_bits(self, y, bits):
        """bits (int32 [..., C/32]) <- [y > 0] of a finished bf16 tensor."""
        c = y.shape[-1]
        self._w(float(y.numel() * y.element_size()), "byte")
        self.call("segk_relu_bits", _p(y), _p(bits), y.numel() // c, c, _stream())
        return bits

    def conv2d_dgrad(self, dy, wd, dx, kh, kw, relu_mask=None, residual=None, scale=1.0, flops=None, colsum=None,
                     relu_mask_bits=None):
        """`colsum` (fp32 [Cin]): column sums of dx from the same pass = BiasAddGrad of the producer layer.
        `relu_mask_bits`: the producer's 1-bit mask (conv2d_fwd's relu_bits) instead of the bf16 tensor `relu_mask`."""
        n, h, w, cout = dy.shape
        cin = dx.shape[3]
        self._w(conv_flops(n, h, w, cin, cout, kh, kw) if flops is None else flops, "flop")
        self._views(tin=dy)
        self.call("segk_conv2d_dgrad", _p(dy), _p(wd), _p(relu_mask), _p(relu_mask_bits), _p(residual), _p(dx), _p(colsum),
                  float(scale), n, h, w, cin, cout, kh, kw, _stream())
        return dx

    def conv2d_wgrad(self, x, dy, dw, kh, kw, accumulate=False, flops=None):
        n, h, w, cin = x.shape
        cout = dy.shape[3]
        self._w(conv_flops(n, h, w, cin, cout, kh, kw) if flops is None else flops, "flop")
        self._views(tin=dy)
        self.call("segk_conv2d_wgrad", _p(x), _p(dy), _p(dw), n, h, w, cin, cout, kh, kw, int(accumulate), _stream())
        return dw

    def deconv2d_fwd(self, x, wk, bias, y, k, s, residual=None, relu=False):
        n, h, w, cin = x.shape
        cout = y.shape[3]
        flags = (EPI_RELU if relu else 0) | (EPI_OUT_F32 if y.dtype == torch.float32 else 0)
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self._views(tout=y)
        self.call("segk_deconv2d_fwd", _p(x), _p(wk), _p(bias), _p(residual), _p(y), n, h, w, cin, cout, k, s, flags,
                  _stream())
        return y

    def deconv2d_dgrad(self, dy, wd, dx, k, s, relu_mask=None, colsum=None):
        n, h, w, cin = dx.shape
        cout = dy.shape[3]
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self._views(tin=dy)
        self.call("segk_deconv2d_dgrad", _p(dy), _p(wd), _p(relu_mask), _p(dx), _p(colsum), n, h, w, cin, cout, k, s, _stream())
        return dx

    def deconv2d_wgrad(self, x, dy, dw, k, s, accumulate=False):
        n, h, w, cin = x.shape
        cout = dy.shape[3]
        self._w(deconv_flops(n, h, w, cin, cout, k, s), "flop")
        self._views(tin=dy)
        self.call("segk_deconv2d_wgrad", _p(x), _p(dy), _p(dw), n, h, w, cin, cout, k, s, int(accumulate), _stream())
        return dw

    # ---- Conv2D 4x4 stride 2 SAME (LidCamNet.py:28-33) on the transposed conv's kernels ---------------
    def conv2d_s2_fwd(self, x, wd, bias, y, relu=True):
        """x [N,2H,2W,Cin] -> y [N,H,W,Cout]; wd from pack_deconv_weights(W_hwio [4,4,Cin,Cout], 2)[1]."""
        n, h, w, cout = y.shape
        cin = x.shape[3]
        self._w(deconv_flops(n, h, w, cout, cin, 4, 2), "flop")
        self.call("segk_conv2d_strided_fwd", _p(x), _p(wd), _p(bias), _p(y), n, h, w, cin, cout, 4, 2, EPI_RELU if relu else 0, _stream())
        return y

    def conv2d_s2_dgrad(self, dy, wk, dx):
        """dx [N,2H,2W,Cin] = conv2d_transpose(dy [N,H,W,Cout], W); wk from pack_deconv_weights(W_hwio, 2)[0]."""
        return self.deconv2d_fwd(dy, wk, None, dx, 4, 2)

    def conv2d_s2_wgrad(self, x, dy, dw, accumulate=False):
        """dw [4,4,Cin,Cout] (HWIO) from x [N,2H,2W,Cin] and dy [N,H,W,Cout]."""
        return self.deconv2d_wgrad(dy, x, dw, 4, 2, accumulate=accumulate)

    # ---- CUDA-core layers for ragged channel counts -------------------------------------
    def conv2d_small_fwd(self, x, w, bias, y, relu=True):
        n, h, wd_, cin = x.shape
        kh, kw, _, cout = w.shape
        self._w(conv_flops(n, h, wd_, cin, cout, kh, kw), "flop")
        self.call("segk_conv2d_small_fwd", _p(x), _dt(x), _p(w), _p(bias), _p(y), n, h, wd_, cin, cout, kh, kw,
                  (EPI_RELU if relu else 0) | (EPI_OUT_F32 if y.dtype == torch.float32 else 0), _stream())
        return y

    def conv2d_small_dgrad(self, dy, w, dx, relu_mask=None, scale=1.0):
        n, h, wd_, cout = dy.shape
        kh, kw, cin, _ = w.shape
        self._w(conv_flops(n, h, wd_, cin, cout, kh, kw), "flop")
        self.call("segk_conv2d_small_dgrad", _p(dy), _p(w), _p(relu_mask), _p(dx), float(scale), n, h, wd_, cin, cout,
                  kh, kw, _stream())
        return dx

    def conv2d_small_wgrad(self, x, dy, dw):
        n, h, wd_, cin = x.shape
        kh, kw, _, cout = dw.shape
        self._w(conv_flops(n, h, wd_, cin, cout, kh, kw), "flop")
        self.call("segk_conv2d_small_wgrad", _p(x), _dt(x), _p(dy), _p(dw), n, h, wd_, cin, cout, kh, kw, _stream())
        return dw

    def deconv2d_small_fwd(self, x, w, bias, y, s, residual=None, relu=False):
        n, h, wd_, cin = x.shape
        k, _, cout, _ = w.shape
        flags = (EPI_RELU if relu else 0) | (EPI_OUT_F32 if y.dtype == torch.float32 else 0)
        self._w(deconv_flops(n, h, wd_, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_small_fwd", _p(x), _p(w), _p(bias), _p(residual), _p(y), n, h, wd_, cin, cout, k, s,
                  flags, _stream())
        return y

    def deconv2d_small_dgrad(self, dy, w, dx, s, relu_mask=None):
        n, h, wd_, cin = dx.shape
        k, _, cout, _ = w.shape
        self._w(deconv_flops(n, h, wd_, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_small_dgrad", _p(dy), int(dy.dtype == torch.float32), _p(w), _p(relu_mask), _p(dx), n,
                  h, wd_, cin, cout, k, s, _stream())
        return dx

    def deconv2d_small_wgrad(self, x, dy, dw, s):
        n, h, wd_, cin = x.shape
        k, _, cout, _ = dw.shape
        self._w(deconv_flops(n, h, wd_, cin, cout, k, s), "flop")
        self.call("segk_deconv2d_small_wgrad", _p(x), _p(dy), int(dy.dtype == torch.float32), _p(dw), n, h, wd_, cin,
                  cout, k, s, _stream())
        return dw

    # ---- HBM-bound kernels ----------------------------------------------------------------
    def conv2d_fwd_pool(self, x, wk, bias, y, pooled, idx, kh, kw, relu=True, pool_only=False):
        """conv_layer + max_pool 2x2 in one launch (pool in the conv epilogue); `pool_only`: y need not be written."""
        n, h, w, cin = x.shape
        cout = y.shape[3]
        self._w(conv_flops(n, h, w, cin, cout, kh, kw), "flop")
        self._views(tin=x, tout=y)
        self.call("segk_conv2d_fwd_pool", _p(x), _p(wk), _p(bias), _p(y), _p(pooled), _p(idx), int(pool_only), n, h, w, cin,
                  cout, kh, kw, EPI_RELU if relu else 0, _stream())
        return pooled, idx

    def maxpool_fwd(self, x, y, idx):
        n, h, w, c = x.shape
        self._w(2.0 * x.numel() + 3.0 * y.numel(), "byte")
        self._views(tin=x)
        self.call("segk_maxpool2x2_fwd", _p(x), _p(y), _p(idx), n, h, w, c, _stream())
        return y, idx

    def maxpool_bwd(self, dy, idx, dx, act=None, residual=None, pooled=None, dbias=None):
        """`act`: pre-pool activation as ReluGrad mask; `pooled`: the pooled tensor as the same mask
        (bit-identical, a quarter of the bytes; not with `residual`); `dbias`: BiasAddGrad of the conv in
        front of the pool from the same pass (pooled mode)."""
        n, h, w, c = dx.shape
        if pooled is not None:
            assert act is None and residual is None
            self._w(2.0 * dx.numel() + 5.0 * dy.numel(), "byte")
            self.call("segk_maxpool2x2_bwd", _p(dy), _p(idx), _p(pooled), 1, 0, _p(dx), _p(dbias), n, h, w, c, _stream())
            return dx
        assert dbias is None
        extra = (2.0 * dx.numel() if act is not None else 0.0) + (2.0 * dx.numel() if residual is not None else 0.0)
        self._w(2.0 * dx.numel() + 3.0 * dy.numel() + extra, "byte")
        pd = _pitch(dx)
        if pd and ((act is not None and _pitch(act) != pd) or (residual is not None and _pitch(residual) != pd)):
            raise ValueError("maxpool_bwd: dx, act and residual must be views of the same geometry")
        self._views(tout=dx)
        self.call("segk_maxpool2x2_bwd", _p(dy), _p(idx), _p(act), 0, _p(residual), _p(dx), 0, n, h, w, c, _stream())
        return dx

    # ---- remaining op families (csrc/opfam.cu) -----------------------------------------
    def avgpool_window_fwd(self, x, y, kh, kw, sh, sw):
        """Avg_Pooling(x, kh, kw, sh, sw, VALID) (utils.py:309)."""
        n, h, w, c = x.shape
        self._w(2.0 * x.numel() + 2.0 * y.numel(), "byte")
        self.call("segk_avgpool_fwd", _p(x), _p(y), n, h, w, c, kh, kw, sh, sw, _stream())
        return y

    def avgpool_window_bwd(self, dy, dx, kh, kw, sh, sw):
        n, h, w, c = dx.shape
        self._w(2.0 * dx.numel() + 2.0 * dy.numel(), "byte")
        self.call("segk_avgpool_bwd", _p(dy), _p(dx), n, h, w, c, kh, kw, sh, sw, _stream())
        return dx

    def maxpool_general_fwd(self, x, y, idx, kh, kw, stride, same=False):
        """Max_Pooling(x, kh, kw, stride, VALID | SAME) (utils.py:306) with first-max indices."""
        n, h, w, c = x.shape
        self._w(2.0 * x.numel() + 3.0 * y.numel(), "byte")
        self.call("segk_maxpool_fwd", _p(x), _p(y), _p(idx), n, h, w, c, kh, kw, stride, int(same), _stream())
        return y, idx

    def maxpool_general_bwd(self, dy, idx, dx, kh, kw, stride, same=False):
        n, h, w, c = dx.shape
        self._w(2.0 * dx.numel() + 3.0 * dy.numel(), "byte")
        self.call("segk_maxpool_bwd", _p(dy), _p(idx), _p(dx), n, h, w, c, kh, kw, stride, int(same), _stream())
        return dx

    def depthwise_conv2d_fwd(self, x, w, bias, y, stride=1, rate=1, relu=False):
        """tf.nn.depthwise_conv2d(x, w [kh,kw,C,1], stride, SAME, rate); w fp32 [kh,kw,C]."""
        n, h, wd, c = x.shape
        kh, kw = w.shape[0], w.shape[1]
        self._w(2.0 * x.numel() + 2.0 * y.numel(), "byte")
        self.call("segk_depthwise_conv2d_fwd", _p(x), _p(w), _p(bias), _p(y), n, h, wd, c, kh, kw, stride, rate,
                  EPI_RELU if relu else 0, _stream())
        return y

    def depthwise_conv2d_dgrad(self, dy, w, dx, stride=1, rate=1):
        n, h, wd, c = dx.shape
        self._w(2.0 * dx.numel() + 2.0 * dy.numel(), "byte")
        self.call("segk_depthwise_conv2d_dgrad", _p(dy), _p(w), _p(dx), n, h, wd, c, w.shape[0], w.shape[1], stride, rate, _stream())
        return dx

    def depthwise_conv2d_wgrad(self, x, dy, dw, stride=1, rate=1, accumulate=False):
        n, h, wd, c = x.shape
        self._w(2.0 * x.numel() * dw.shape[0] * dw.shape[1] + 2.0 * dy.numel(), "byte")
        self.call("segk_depthwise_conv2d_wgrad", _p(x), _p(dy), _p(dw), n, h, wd, c, dw.shape[0], dw.shape[1], stride, rate,
                  int(accumulate), _stream())
        return dw

    def activation_fwd(self, x, y, kind):
        """kind: 'sigmoid' | 'swish'."""
        self.call("segk_activation_fwd", _p(x), _p(y), x.numel(), {"sigmoid": 0, "swish": 1}[kind], _stream())
        return y

    def activation_bwd(self, x, dy, dx, kind):
        self.call("segk_activation_bwd", _p(x), _p(dy), _p(dx), x.numel(), {"sigmoid": 0, "swish": 1}[kind], _stream())
        return dx

    def channel_scale_fwd(self, x, s, y):
        """y[n,h,w,c] = x[n,h,w,c] * s[n,c] (squeeze-excite multiply)."""
        n, h, w, c = x.shape
        self.call("segk_channel_scale_fwd", _p(x), _p(s), _p(y), n, h * w, c, _stream())
        return y

    def channel_scale_bwd(self, x, s, dy, dx, ds):
        n, h, w, c = x.shape
        self.call("segk_channel_scale_bwd", _p(x), _p(s), _p(dy), _p(dx), _p(ds), n, h * w, c, _stream())
        return dx, ds

    # ---- shared-helper layers (utils.py): BN-affine folding, concat ---------------------------------
    def scale_columns(self, w, scale, mult, out=None):
        c = w.shape[-1]
        if out is None:
            out = torch.empty_like(w)
        self.call("segk_scale_columns", _p(w), _p(scale), float(mult), _p(out), w.numel() // c, c, _stream())
        return out

    def bn_unfold_workspace(self, max_c, device):
        return torch.empty(2 * self.ctx.sm_count * max(int(max_c), 32), dtype=torch.float32, device=device)

    def bn_unfold_grads(self, gw, w, gamma, mult, dgamma, workspace):
        """gw: gradient of the BN-folded weights -> (in place) gradient of the unfolded weights; dgamma from gw . w."""
        c = gw.shape[-1]
        self._w(12.0 * gw.numel(), "byte")
        self.call("segk_bn_unfold_grads", _p(gw), _p(w), _p(gamma), float(mult), _p(dgamma), _p(workspace),
                  workspace.numel() * workspace.element_size(), gw.numel() // c, c, _stream())
        return dgamma

    def bn_gamma_grad(self, dz, y, beta, gamma, dgamma, workspace, dbeta=None):
        c = dz.shape[-1]
        self._w(4.0 * dz.numel(), "byte")
        self.call("segk_bn_gamma_grad", _p(dz), _p(y), _p(beta), _p(gamma), _p(dgamma), _p(dbeta), _p(workspace),
                  workspace.numel() * workspace.element_size(), dz.numel() // c, c, _stream())
        return dgamma

    def bn_grads_f32(self, dz, y, beta, gamma, dgamma, dbeta, workspace):
        c = dz.shape[-1]
        self._w(8.0 * dz.numel(), "byte")
        self.call("segk_bn_grads_f32", _p(dz), _p(y), _p(beta), _p(gamma), _p(dgamma), _p(dbeta), _p(workspace),
                  workspace.numel() * workspace.element_size(), dz.numel() // c, c, _stream())

    # ---- dense-block builders (FCDenseNet.py): strided channel views, physical <-> logical remaps ----
    def bn_act_fwd(self, x, c, y, scale, shift, relu=True, drop=None):
        """x, y: [..., ld] bf16 buffers; the first c channels of every row are processed.  `drop` = (keep_prob, seed,
        u8 mask or None): the Dropout in front of the BN applied to x on the fly."""
        rows = x.numel() // x.shape[-1]
        self._w(4.0 * rows * c, "byte")
        keep, seed, dmask = drop if drop is not None else (1.0, 0, None)
        self.call("segk_bn_act_fwd", _p(x), x.shape[-1], _p(y), y.shape[-1], _p(scale), _p(shift), rows, c, int(relu),
                  _p(dmask), float(keep), int(seed), _stream())
        return y

    def bn_act_bwd(self, dy, y, x, dx, c, scale, dscale, dshift, workspace, relu=True, accumulate=True, drop=None):
        rows = x.numel() // x.shape[-1]
        self._w((8.0 + (4.0 if accumulate else 2.0)) * rows * c, "byte")
        keep, seed, dmask = drop if drop is not None else (1.0, 0, None)
        self.call("segk_bn_act_bwd", _p(dy), _p(y), y.shape[-1], _p(x), _p(dx), x.shape[-1], _p(scale), _p(dscale), _p(dshift),
                  _p(workspace), workspace.numel() * workspace.element_size(), rows, c, int(relu), int(accumulate),
                  _p(dmask), float(keep), int(seed), _stream())

    def bn_act_bwd_workspace(self, c, device):
        n = int(self.ctx.c.segk_bn_act_bwd_workspace_bytes(self.ctx.h, int(c)))
        return torch.empty(n, dtype=torch.uint8, device=device)

    def avgpool_fwd(self, x, c, y):
        n, h, w, ldx = x.shape
        self._w(2.5 * n * h * w * c, "byte")
        self.call("segk_avgpool2x2_fwd", _p(x), ldx, _p(y), y.shape[-1], n, h, w, c, _stream())
        return y

    def avgpool_bwd(self, dy, c, dx):
        n, h, w, lddx = dx.shape
        self._w(2.5 * n * h * w * c, "byte")
        self.call("segk_avgpool2x2_bwd", _p(dy), dy.shape[-1], _p(dx), lddx, n, h, w, c, _stream())
        return dx

    def remap_weights(self, w, wp, amap=None, bmap=None, to_phys=True):
        """w logical [T..., A, B] fp32 <-> wp physical [T..., Ap, Bp]."""
        a, b = w.shape[-2], w.shape[-1]
        ap, bp = wp.shape[-2], wp.shape[-1]
        t = w.numel() // (a * b)
        self.call("segk_remap_weights", _p(w), _p(wp), t, a, b, ap, bp, _p(amap), _p(bmap), int(to_phys), _stream())

    def gather_f32(self, src, cmap, dst, mul=1.0, add=0.0):
        self.call("segk_gather_f32", _p(src), _p(cmap), _p(dst), dst.numel(), float(mul), float(add), _stream())
        return dst

    def scatter_f32(self, src, cmap, dst, mul=1.0):
        self.call("segk_scatter_f32", _p(src), _p(cmap), _p(dst), src.numel(), float(mul), _stream())
        return dst

    # ---- remaining op families of utils.py: atrous conv, align-corners bilinear resize, global average pool ----
    def atrous_conv2d_fwd(self, x, wk, bias, y, k, rate, relu=False, residual=None):
        n, h, w, cin = x.shape
        cout = y.shape[3]
        flags = (EPI_RELU if relu else 0) | (EPI_OUT_F32 if y.dtype == torch.float32 else 0)
        self._w(2.0 * n * h * w * k * k * cin * cout, "flop")
        self.call("segk_atrous_conv2d_fwd", _p(x), _p(wk), _p(bias), _p(residual), _p(y), n, h, w, cin, cout, k, k, int(rate),
                  flags, _stream())
        return y

    def atrous_conv2d_dgrad(self, dy, wd, dx, k, rate, relu_mask=None, residual=None, scale=1.0, colsum=None):
        n, h, w, cout = dy.shape
        cin = dx.shape[3]
        self._w(2.0 * n * h * w * k * k * cin * cout, "flop")
        self.call("segk_atrous_conv2d_dgrad", _p(dy), _p(wd), _p(relu_mask), _p(residual), _p(dx), _p(colsum), float(scale), n, h,
                  w, cin, cout, k, k, int(rate), _stream())
        return dx

    def atrous_conv2d_wgrad(self, x, dy, dw, k, rate, accumulate=False):
        n, h, w, cin = x.shape
        cout = dy.shape[3]
        self._w(2.0 * n * h * w * k * k * cin * cout, "flop")
        self.call("segk_atrous_conv2d_wgrad", _p(x), _p(dy), _p(dw), n, h, w, cin, cout, k, k, int(rate), int(accumulate), _stream())
        return dw

    def resize_bilinear_fwd(self, x, y):
        n, h, w, c = x.shape
        self.call("segk_resize_bilinear_fwd", _p(x), _p(y), n, h, w, y.shape[1], y.shape[2], c, _stream())
        return y

    def resize_bilinear_bwd(self, dy, dx):
        n, h, w, c = dx.shape
        self.call("segk_resize_bilinear_bwd", _p(dy), _p(dx), n, h, w, dy.shape[1], dy.shape[2], c, _stream())
        return dx

    def global_avgpool_fwd(self, x, y):
        n, h, w, c = x.shape
        self.call("segk_global_avgpool_fwd", _p(x), _p(y), n, h, w, c, _stream())
        return y

    def global_avgpool_bwd(self, dy, dx):
        n, h, w, c = dx.shape
        self.call("segk_global_avgpool_bwd", _p(dy), _p(dx), n, h, w, c, _stream())
        return dx

    def global_maxpool_fwd(self, x, y, count):
        n, h, w, c = x.shape
        self.call("segk_global_maxpool_fwd", _p(x), _p(y), _p(count), n, h, w, c, _stream())
        return y

    def global_maxpool_bwd(self, dy, x, y, count, dx):
        n, h, w, c = x.shape
        self.call("segk_global_maxpool_bwd", _p(dy), _p(x), _p(y), _p(count), _p(dx), n, h, w, c, _stream())
        return dx

    def zero_pad(self, x, y, pad, crop=False):
        """crop=False: y [N,H+2p,W+2p,C] <- x [N,H,W,C] (Zero_Padding); crop=True: its gradient, y [N,H,W,C] <- centre of x."""
        n, h, w, c = (y if crop else x).shape
        self.call("segk_zero_pad", _p(x), _p(y), n, h, w, c, int(pad), int(crop), _stream())
        return y

    def channel_copy(self, src, coff_src, dst, coff_dst, c, mask=None, accumulate=False, drop=None):
        """`drop` = (side, keep_prob, seed, u8 mask or None): dropout of the copied values on the fly, pattern indexed by
        the source (side 1) or destination (side 2) tensor."""
        rows = src.numel() // src.shape[-1]
        self._w(4.0 * rows * c, "byte")
        side, keep, seed, dmask = drop if drop is not None else (0, 1.0, 0, None)
        self.call("segk_channel_copy", _p(src), src.shape[-1], coff_src, _p(dst), dst.shape[-1], coff_dst, _p(mask),
                  int(accumulate), rows, c, int(side), _p(dmask), float(keep), int(seed), _stream())
        return dst

    def dropout(self, x, y, keep_prob, seed, mask=None):
        self.call("segk_dropout", _p(x), _p(y), _p(mask), x.numel(), float(keep_prob), int(seed), _stream())
        return y

    def xent_workspace(self, npix, device):
        nbytes = int(self.ctx.c.segk_xent_workspace_bytes(int(npix)))
        return torch.empty(nbytes, dtype=torch.uint8, device=device)

    def softmax_xent(self, logits, labels, dlogits, pred, loss_sum, cm, workspace, grad_scale):
        c = logits.shape[-1]
        npix = logits.numel() // c
        self._w(npix * (4.0 * c + 1.0 + (4.0 * c if dlogits is not None else 0.0)), "byte")
        self.call("segk_softmax_xent_fwd_bwd", _p(logits), _p(labels), _p(dlogits), _p(pred), _p(loss_sum), _p(cm),
                  _p(workspace), npix, c, float(grad_scale), _stream())

    def softmax_infer(self, logits, prob=None, mask=None, argmax=None):
        c = logits.shape[-1]
        self.call("segk_softmax_infer", _p(logits), _p(prob), _p(mask), _p(argmax), logits.numel() // c, c, _stream())

    def onehot_to_ids(self, onehot, ids):
        """[..., C] one-hot annotation (bool / u8 / f32, FCN.py:313) -> u8 class ids."""
        c = onehot.shape[-1]
        if onehot.dtype == torch.bool:
            onehot = onehot.view(torch.uint8)
        self.call("segk_onehot_to_ids", _p(onehot), _dt(onehot), _p(ids), onehot.numel() // c, c, _stream())
        return ids

    def overlay_mask(self, image, mask, out=None, color=(0, 255, 0, 127)):
        """paste_mask of FCN.py:203-211 on the GPU (u8 NHWC image, u8 mask)."""
        if out is None:
            out = torch.empty_like(image)
        self.call("segk_overlay_mask", _p(image), _p(mask), _p(out), mask.numel(), image.shape[-1], int(color[0]),
                  int(color[1]), int(color[2]), int(color[3]), _stream())
        return out

    def confusion_matrix(self, gt, pred, cm):
        self.call("segk_confusion_matrix", _p(gt), _p(pred), _p(cm), gt.numel(), _stream())
        return cm

    def adam_step(self, p, m, v, g, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
        self._w(28.0 * p.numel(), "byte")
        self.call("segk_adam_step", _p(p), _p(m), _p(v), _p(g), p.numel(), float(lr_t), float(beta1), float(beta2),
                  float(eps), float(grad_scale), _stream())

    def adam_step_ranges(self, p, m, v, g, ranges, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
        """ApplyAdam on the (offset, length) element ranges of the flat arenas, one launch."""
        import ctypes
        n = len(ranges)
        offs = (ctypes.c_int64 * n)(*[int(o) for o, _ in ranges])
        lens = (ctypes.c_int64 * n)(*[int(l) for _, l in ranges])
        self._w(28.0 * sum(int(l) for _, l in ranges), "byte")
        self.call("segk_adam_step_ranges", _p(p), _p(m), _p(v), _p(g), ctypes.addressof(offs), ctypes.addressof(lens), n,
                  float(lr_t), float(beta1), float(beta2), float(eps), float(grad_scale), _stream())

    def adam_pack_conv_weights(self, p, m, v, g, wk, wd, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, col_scale=None,
                               col_mult=1.0):
        """p, m, v, g: fp32 views [kh,kw,Cin,Cout] of one conv weight; wk / wd its bf16 kernel layouts.  `col_scale` (fp32
        [Cout]) * col_mult: the folded BN scale the packed copies carry (Conv2D_Block, utils.py:186-208)."""
        kh, kw, cin, cout = p.shape
        self._w(32.0 * p.numel(), "byte")
        self.call("segk_adam_pack_conv_weights", _p(p), _p(m), _p(v), _p(g), _p(wk), _p(wd), _p(col_scale), float(col_mult), kh, kw,
                  cin, cout, float(lr_t), float(beta1), float(beta2), float(eps), float(grad_scale), _stream())

    def momentum_step(self, p, a, g, lr, mu, grad_scale=1.0):
        self.call("segk_momentum_step", _p(p), _p(a), _p(g), p.numel(), float(lr), float(mu), float(grad_scale),
                  _stream())

    def cast_to_bf16(self, x, y):
        self.call("segk_cast_to_bf16", _p(x), _dt(x), _p(y), x.numel(), _stream())
        return y

    def bias_grad(self, dy, db):
        c = dy.shape[-1]
        self._w(float(dy.numel() * dy.element_size()), "byte")
        self._views(tin=dy)
        self.call("segk_bias_grad", _p(dy), int(dy.dtype == torch.float32), _p(db), dy.numel() // c, c, _stream())
        return db
