"""FCN-8s CPU fp32 oracle.  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Restates `Network/model/FCN.py:52-107` (graph), `:117-171` (layer helpers), `:334`
(loss), `:338-340` (Adam) with the TF op semantics in oracle/tf_ops.py.  The backward pass
is torch autograd over those ops (TF's `optimizer.minimize` synthesises the same graph).

PARITY UNPINNED: no reference test or golden vector exists for this path.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as T

# (name, kh, kw, cin, cout) in reference creation order, FCN.py:52-86.  cin=None -> input.
ENCODER = [
    ("conv1_1", 3, None, 64), ("conv1_2", 3, 64, 64), "pool1",
    ("conv2_1", 3, 64, 128), ("conv2_2", 3, 128, 128), "pool2",
    ("conv3_1", 3, 128, 256), ("conv3_2", 3, 256, 256), ("conv3_3", 3, 256, 256), "pool3",
    ("conv4_1", 3, 256, 512), ("conv4_2", 3, 512, 512), ("conv4_3", 3, 512, 512),
    ("conv4_4", 3, 512, 512), "pool4",
    ("conv5_1", 3, 512, 512), ("conv5_2", 3, 512, 512), ("conv5_3", 3, 512, 512), "pool5",
]


def variable_shapes(cin: int = 3, ncls: int = 2, fc: int = 4096):
    """Ordered {name: shape} of the 40 variables, creation order of FCN.py:52-107."""
    shapes = OrderedDict()
    for item in ENCODER:
        if isinstance(item, str):
            continue
        name, k, ci, co = item
        ci = cin if ci is None else ci
        shapes[f"{name}/weights"] = (k, k, ci, co)
        shapes[f"{name}/biases"] = (co,)
    shapes["conv6/weights"] = (7, 7, 512, fc)
    shapes["conv6/biases"] = (fc,)
    shapes["conv7/weights"] = (1, 1, fc, fc)
    shapes["conv7/biases"] = (fc,)
    shapes["conv8/weights"] = (1, 1, fc, ncls)
    shapes["conv8/biases"] = (ncls,)
    # deconv_layer: W [fh, fw, shape[3] (=Cout), num_filters (=Cin)]  (FCN.py:143)
    shapes["conv_t1/weights"] = (4, 4, 512, ncls)
    shapes["conv_t1/biases"] = (512,)
    shapes["conv_t2/weights"] = (4, 4, 256, 512)
    shapes["conv_t2/biases"] = (256,)
    shapes["conv_t3/weights"] = (16, 16, ncls, 256)
    shapes["conv_t3/bias"] = (ncls,)   # singular, FCN.py:103
    return shapes


def init_variables(cin=3, ncls=2, fc=4096, seed=1234, init="ref"):
    """Deterministic variables: one default_rng(seed) stream consumed in
    creation order.  init='ref' -> N(0,0.01^2) weights (FCN.py:125), zero biases (:127);
    init='he' -> std sqrt(2/fan_in) so every layer is numerically visible."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(cin, ncls, fc).items():
        if name.endswith("weights"):
            z = rng.standard_normal(shape, dtype=np.float32)
            if init == "ref":
                std = 0.01
            else:
                if name.startswith("conv_t"):
                    kh, kw, co, ci = shape
                    s = {4: 2, 16: 8}[kh]
                    fan_in = (kh // s) * (kw // s) * ci
                else:
                    kh, kw, ci, co = shape
                    fan_in = kh * kw * ci
                std = float(np.sqrt(2.0 / fan_in))
            out[name] = (z * np.float32(std)).astype(np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


def synthetic_batch(n, h, w, cin=3, seed=0, road_shaped=False):
    """Synthetic input: raw 0..255 pixels, class-id labels (1 = road)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, (n, h, w, cin), dtype=np.uint8)
    if road_shaped:
        yy, xx = np.mgrid[0:h, 0:w]
        half = (yy - h // 2) * (w // 2) // max(h // 2, 1)
        lab = ((yy >= h // 2) & (np.abs(xx - w // 2) <= half)).astype(np.uint8)
        lab = np.broadcast_to(lab, (n, h, w)).copy()
    else:
        lab = rng.integers(0, 2, (n, h, w), dtype=np.uint8)
    return x, lab


class _RoundBoth(torch.autograd.Function):
    """bf16 storage point: value rounded in forward, gradient rounded in backward (the CUDA path
    stores both activations and activation gradients as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return T.to_bf16_grid(x)

    @staticmethod
    def backward(ctx, g):
        return T.to_bf16_grid(g)


class FCN8sOracle:
    """Forward/backward/Adam of the FCN.py graph on CPU.

    bf16_storage=True rounds weights and every stored activation to the bf16 grid (fp32
    arithmetic in between), mirroring where the CUDA path stores bf16; logits stay fp32.
    """

    def __init__(self, variables, ncls=2, bf16_storage=False, threads=None, bf16_grads=False):
        if threads:
            torch.set_num_threads(threads)
        self.ncls = ncls
        self.bf16 = bf16_storage
        self.bf16_grads = bf16_grads      # also round activation gradients at the storage points
        self.vars = OrderedDict((k, torch.tensor(v, dtype=torch.float32, requires_grad=True))
                                for k, v in variables.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
        self.t = 0
        self.acts = OrderedDict()

    def _q(self, x, weight=False):
        if not self.bf16:
            return x
        if self.bf16_grads and not weight:
            return _RoundBoth.apply(x)
        # straight-through rounding so autograd matches "store bf16, compute fp32"
        return x + (T.to_bf16_grid(x.detach()) - x.detach())

    def _w(self, name):
        return self._q(self.vars[name], weight=True)

    def _conv(self, x, name, keep=True):
        # conv_layer, FCN.py:117-136
        z = T.bias_add(T.conv2d_same(x, self._w(f"{name}/weights")), self.vars[f"{name}/biases"])
        a = self._q(T.relu(z))
        if keep:
            self.acts[name] = a
        return a

    def _deconv(self, x, name, out_hw, stride, bias_name="biases"):
        # deconv_layer, FCN.py:138-159 ; conv_t3 inline FCN.py:101-107
        y = T.conv2d_transpose_same(x, self._w(f"{name}/weights"), out_hw, stride)
        return T.bias_add(y, self.vars[f"{name}/{bias_name}"])

    def forward(self, x_u8_or_f32, keep_prob=1.0, masks=None):
        """x: [N,H,W,Cin] raw 0..255.  Returns (pred int64 [N,H,W,1], logits f32 [N,H,W,C])."""
        x = torch.as_tensor(np.asarray(x_u8_or_f32), dtype=torch.float32)
        self.acts.clear()
        self.acts["input"] = x
        h = x
        for item in ENCODER:
            if isinstance(item, str):
                h = T.max_pool_2x2(h)
                self.acts[item] = h
            else:
                h = self._conv(h, item[0])
        pool3, pool4 = self.acts["pool3"], self.acts["pool4"]
        masks = masks or {}
        h = self._conv(h, "conv6")
        h = self._q(T.dropout(h, keep_prob, masks.get("dropout6")))
        self.acts["dropout6"] = h
        h = self._conv(h, "conv7")
        h = self._q(T.dropout(h, keep_prob, masks.get("dropout7")))
        self.acts["dropout7"] = h
        conv8 = self._conv(h, "conv8")                                   # ReLU'd, FCN.py:86
        t1 = self._deconv(conv8, "conv_t1", pool4.shape[1:3], 2)
        fuse_1 = self._q(t1 + pool4)                                     # FCN.py:92
        self.acts["fuse_1"] = fuse_1
        t2 = self._deconv(fuse_1, "conv_t2", pool3.shape[1:3], 2)
        fuse_2 = self._q(t2 + pool3)                                     # FCN.py:96
        self.acts["fuse_2"] = fuse_2
        logits = self._deconv(fuse_2, "conv_t3", x.shape[1:3], 8, bias_name="bias")
        self.acts["logits"] = logits
        pred = T.argmax_last(logits).unsqueeze(3)                        # FCN.py:111-114
        return pred, logits

    def loss(self, logits, labels_u8):
        """reduce_mean(softmax_cross_entropy_with_logits) with one-hot labels (FCN.py:334).
        labels: [N,H,W] class ids (channel 0 = background, 1 = road, FCN.py:195-201)."""
        lab = torch.as_tensor(np.asarray(labels_u8), dtype=torch.int64)
        onehot = torch.nn.functional.one_hot(lab, self.ncls).to(torch.float32)
        return T.softmax_cross_entropy_with_logits(logits, onehot).mean()

    def loss_and_grads(self, x, labels, keep_prob=1.0, masks=None, retain=()):
        for v in self.vars.values():
            v.grad = None
        _, logits = self.forward(x, keep_prob, masks)
        for k in retain:
            self.acts[k].retain_grad()
        loss = self.loss(logits, labels)
        loss.backward()
        grads = OrderedDict((k, v.grad.detach().clone()) for k, v in self.vars.items())
        return float(loss.detach()), logits.detach(), grads

    def train_step(self, x, labels, keep_prob=1.0, masks=None, lr=1e-4):
        """One sess.run(train_step) (FCN.py:398): fwd + bwd + TF-Adam on all 40 variables."""
        loss, logits, grads = self.loss_and_grads(x, labels, keep_prob, masks)
        self.t += 1
        with torch.no_grad():
            for k, p in self.vars.items():
                T.adam_tf_step(p, self.m[k], self.v[k], grads[k], self.t, lr=lr)
        return loss, logits, grads


def default_threads():
    return len(os.sched_getaffinity(0))
