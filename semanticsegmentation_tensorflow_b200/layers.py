"""The reference's inline layer helpers (`Network/model/FCN.py:117-171`) as eager functions over the B200
kernels, with the reference's signatures:

    conv_layer(x, num_filters, name, filter_height=3, filter_width=3, stride=1, padding='SAME')   FCN.py:117
    deconv_layer(x, shape, num_filters, name, output_shape, filter_height=4, filter_width=4,
                 stride=2, padding='SAME')                                                         FCN.py:138
    max_pool(x, name, filter_height=2, filter_width=2, stride=2, padding='VALID')                  FCN.py:161
    dropout(x, keep_prob)                                                                          FCN.py:165
    fuse(x1, x2, name)                                                                             FCN.py:169

The TF-1.x graph / session split collapses: each call runs its kernel on the current CUDA stream and returns a
device tensor (NHWC bf16; the image itself may be u8).  `tf.get_variable(..., reuse=AUTO_REUSE)` becomes a
`VariableStore`: `<name>/weights` is created N(0, 0.01^2) in the reference's layout on first use and reused
afterwards (FCN.py:123-127,141-145).  These helpers compose forward passes; training goes through
`FCN(...)` + `AdamOptimizer(...).minimize(net)`, which plans the backward schedule and buffers.
Unsupported arguments raise (no fallback): stride-1 SAME convs, k = 2*stride transposed convs, 2x2/s2 VALID pools.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .ops import Ops
from .plan import Layer


class VariableStore:
    """tf.get_variable with reuse=AUTO_REUSE: fp32 masters under the reference's names and layouts, plus
    the bf16 kernel-layout shadows the tensor-core kernels read (rebuilt when a master is assigned)."""

    def __init__(self, device=None, seed=1234):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.ops = Ops(self.device)
        self.vars = OrderedDict()
        self._packed = {}
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        self.dropout_calls = 0

    def get_variable(self, name, shape, init_std=None):
        v = self.vars.get(name)
        if v is None:
            if init_std is None:
                v = torch.zeros(shape, dtype=torch.float32, device=self.device)          # constant_initializer(0.0)
            else:                                                                         # random_normal_initializer(0, std)
                v = torch.randn(shape, generator=self._gen, device=self.device, dtype=torch.float32) * init_std
            self.vars[name] = v
        elif tuple(v.shape) != tuple(shape):
            raise ValueError(f"variable {name} exists with shape {tuple(v.shape)}, requested {tuple(shape)}")
        return v

    def assign(self, values):
        for name, arr in values.items():
            t = torch.as_tensor(arr, dtype=torch.float32).to(self.device)
            self.vars[name] = t.contiguous()
            self._packed.pop(name.rsplit("/", 1)[0], None)

    def packed(self, scope, make):
        if scope not in self._packed:
            self._packed[scope] = make()
        return self._packed[scope]


_DEFAULT = {}


def default_store(device=None) -> VariableStore:
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev not in _DEFAULT:
        _DEFAULT[dev] = VariableStore(dev)
    return _DEFAULT[dev]


def _store(x, store):
    return store if store is not None else default_store(x.device)


def conv_layer(x, num_filters, name, filter_height=3, filter_width=3, stride=1, padding="SAME", store=None):
    """relu(conv2d(x, W, stride, SAME) + b) with variables `<name>/weights` HWIO, `<name>/biases` (FCN.py:117-136)."""
    if stride != 1 or padding != "SAME" or filter_height != filter_width:
        raise ValueError(f"conv_layer {name}: only stride 1, SAME, square filters are built (no fallback)")
    st = _store(x, store)
    ops, k, cin = st.ops, filter_height, x.shape[3]
    w = st.get_variable(f"{name}/weights", (k, k, cin, num_filters), 0.01)
    b = st.get_variable(f"{name}/biases", (num_filters,))
    n, h, wd = x.shape[:3]
    route = Layer(name, "conv", k, cin, num_filters).path
    y = torch.empty((n, h, wd, num_filters), dtype=torch.bfloat16, device=x.device)
    if route == "tc":
        wk, _ = st.packed(name, lambda: ops.pack_conv_weights(w))
        ops.conv2d_fwd(x, wk, b, y, k, k, relu=True)
    elif route == "first":
        wk = st.packed(name, lambda: ops.pack_im2col_weights(w))
        ops.conv2d_first_fwd(x, wk, b, y, k, k, relu=True)
    elif route == "im2col":
        wk = st.packed(name, lambda: ops.pack_im2col_weights(w))
        P1 = ops.im2col_k64(x, torch.empty((n, h, wd, 64), dtype=torch.bfloat16, device=x.device), k, k)
        ops.conv2d_fwd(P1, wk, b, y, 1, 1, relu=True)
    else:
        ops.conv2d_small_fwd(x, w, b, y, relu=True)
    return y


def deconv_layer(x, shape, num_filters, name, output_shape, filter_height=4, filter_width=4, stride=2, padding="SAME",
                 store=None):
    """conv2d_transpose(x, W[fh, fw, shape[3], num_filters], output_shape, stride, SAME) + b (FCN.py:138-159).
    `shape`: the shape whose channel count the output takes (the skip tensor's, FCN.py:90,94); `num_filters` is
    the INPUT channel count, as in the reference; `output_shape` None = twice the input size."""
    if padding != "SAME" or filter_height != filter_width or filter_height != 2 * stride:
        raise ValueError(f"deconv_layer {name}: only k = 2*stride, SAME is built (no fallback)")
    st = _store(x, store)
    ops, k = st.ops, filter_height
    cout = int(shape[3])
    n, h, wd, cin = x.shape
    if cin != num_filters:
        raise ValueError(f"deconv_layer {name}: num_filters {num_filters} != input channels {cin}")
    if output_shape is not None and tuple(int(v) for v in output_shape) != (n, h * stride, wd * stride, cout):
        raise ValueError(f"deconv_layer {name}: output_shape {tuple(output_shape)} != {(n, h * stride, wd * stride, cout)}")
    w = st.get_variable(f"{name}/weights", (k, k, cout, cin), 0.01)
    b = st.get_variable(f"{name}/biases", (cout,))
    route = Layer(name, "deconv", k, cin, cout, stride=stride).path
    y = torch.empty((n, h * stride, wd * stride, cout), dtype=torch.bfloat16, device=x.device)
    if route == "tc":
        wk, _ = st.packed(name, lambda: ops.pack_deconv_weights(w, stride))
        ops.deconv2d_fwd(x, wk, b, y, k, stride)
    elif route == "packed":
        bf, _ = st.packed(name, lambda: ops.pack_deconv_packed(w, stride))
        y = torch.empty((n, h * stride, wd * stride, cout), dtype=torch.float32, device=x.device)    # fp32 logits
        ops.deconv2d_packed_fwd(x, bf, b, y, k, stride)
    elif route == "patch":
        e = k * k * cout
        wk, _ = st.packed(name, lambda: ops.pack_matrix(w.view(1, e, cin)))
        yp = ops.conv2d_fwd(x, wk, None, torch.empty((n, h, wd, e), dtype=torch.float32, device=x.device), 1, 1, relu=False)
        ops.deconv_col2im(yp, b, y, k, stride)
    else:
        ops.deconv2d_small_fwd(x, w, b, y, stride)
    return y


def max_pool(x, name, filter_height=2, filter_width=2, stride=2, padding="VALID", store=None):
    """tf.nn.max_pool 2x2 / stride 2 / VALID (FCN.py:161-163)."""
    if (filter_height, filter_width, stride, padding) != (2, 2, 2, "VALID"):
        raise ValueError(f"max_pool {name}: only 2x2 / stride 2 / VALID is built (no fallback)")
    st = _store(x, store)
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
    st.ops.maxpool_fwd(x, y, torch.empty(y.shape, dtype=torch.uint8, device=x.device))
    return y


def dropout(x, keep_prob, store=None, seed=None):
    """tf.nn.dropout(x, keep_prob) (FCN.py:165-167); keep_prob 1.0 is the identity.  Philox4x32-10 stream,
    a fresh seed per call unless one is given."""
    keep_prob = float(keep_prob)
    if keep_prob >= 1.0:
        return x
    st = _store(x, store)
    if seed is None:
        st.dropout_calls += 1
        seed = 0x5E6B0000 + st.dropout_calls
    return st.ops.dropout(x, torch.empty_like(x), keep_prob, seed)


def fuse(x1, x2, name, store=None):
    """tf.add(x1, x2) skip connection (FCN.py:169-171)."""
    if x1.shape != x2.shape or x1.dtype != torch.bfloat16 or x2.dtype != torch.bfloat16:
        raise ValueError(f"fuse {name}: needs two bf16 tensors of one shape")
    st = _store(x1, store)
    out = x2.clone()
    return st.ops.channel_copy(x1, 0, out, 0, x1.shape[3], accumulate=True)
