"""CPU fp32 oracle for the encoder-decoder builders made from the reference's shared helpers
(`Network/utils/utils.py:164-333`).  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

PARITY UNPINNED (no reference tests / fixtures; TensorFlow not installable).

U-Net: not a file of the reference.  Definition restated here independently of
the product code: SegNet's VGG16-BN encoder (`SegNet.py:30-51`) with ReLU on, and the decoder idiom
of `FCDenseNet.py:141-160` — `Deconv2D_Block(4x4, s2)` -> `Concat([up, skip])` -> `Conv2D_Block`s
with SegNet's decoder widths (`SegNet.py:57-81`) — ending in FCDenseNet's 1x1 `final_conv` (`:157`)."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as T
from .fcn_oracle import _RoundBoth

BN_EPS = 1e-3   # tf.layers.batch_normalization default; training=False, moving stats never updated (utils.py:300-301)


def unet_spec(num_classes=2):
    """[(name, kind, inputs, k, cout)]; every 3x3 Conv2D_Block has batch_normalization=True, relu=True."""
    S = []
    def conv(name, src, co):
        S.append((name, "conv_bn_relu", [src], 3, co)); return name
    def pool(name, src):
        S.append((name, "pool", [src], 2, 0)); return name
    def up(name, src, co):
        S.append((name, "deconv", [src], 4, co)); return name
    def cat(name, a, b):
        S.append((name, "concat", [a, b], 0, 0)); return name
    c1 = conv("conv1", "input", 64); c2 = conv("conv2", c1, 64); p1 = pool("pool1", c2)
    c3 = conv("conv3", p1, 128); c4 = conv("conv4", c3, 128); p2 = pool("pool2", c4)
    c5 = conv("conv5", p2, 256); c6 = conv("conv6", c5, 256); c7 = conv("conv7", c6, 256); p3 = pool("pool3", c7)
    c8 = conv("conv8", p3, 512); c9 = conv("conv9", c8, 512); c10 = conv("conv10", c9, 512); p4 = pool("pool4", c10)
    c11 = conv("conv11", p4, 512); c12 = conv("conv12", c11, 512); c13 = conv("conv13", c12, 512); p5 = pool("pool5", c13)
    x = cat("concat1", up("unpool1", p5, 512), c13)
    x = conv("conv14", x, 512); x = conv("conv15", x, 512); x = conv("conv16", x, 512)
    x = cat("concat2", up("unpool2", x, 512), c10)
    x = conv("conv17", x, 512); x = conv("conv18", x, 512); x = conv("conv19", x, 256)
    x = cat("concat3", up("unpool3", x, 256), c7)
    x = conv("conv20", x, 256); x = conv("conv21", x, 256); x = conv("conv22", x, 128)
    x = cat("concat4", up("unpool4", x, 128), c4)
    x = conv("conv23", x, 128); x = conv("conv24", x, 64)
    x = cat("concat5", up("unpool5", x, 64), c2)
    x = conv("conv25", x, 64)
    S.append(("final_conv", "conv", [x], 1, num_classes))
    return S


def segnet_spec(num_classes=2):
    """The reference's SegNet exactly as written (`SegNet.py:28-87`): every `Conv2D_Block(...,
    batch_normalization=True)` keeps the helper's default `relu=False` (`utils.py:194`) -- the network has
    no ReLU --, `Deconv2D_Block` 4x4 s2 "unpool" layers without skip connections, and a 3x3
    `Conv2D_Layer` to `num_classes` followed by `Batch_Normalization` (`SegNet.py:80-81`)."""
    S = []
    def conv(name, src, co):
        S.append((name, "conv_bn", [src], 3, co)); return name
    def pool(name, src):
        S.append((name, "pool", [src], 2, 0)); return name
    def up(name, src, co):
        S.append((name, "deconv", [src], 4, co)); return name
    x = conv("conv1", "input", 64); x = conv("conv2", x, 64); x = pool("pool1", x)
    x = conv("conv3", x, 128); x = conv("conv4", x, 128); x = pool("pool2", x)
    x = conv("conv5", x, 256); x = conv("conv6", x, 256); x = conv("conv7", x, 256); x = pool("pool3", x)
    x = conv("conv8", x, 512); x = conv("conv9", x, 512); x = conv("conv10", x, 512); x = pool("pool4", x)
    x = conv("conv11", x, 512); x = conv("conv12", x, 512); x = conv("conv13", x, 512); x = pool("pool5", x)
    x = up("unpool1", x, 512); x = conv("conv14", x, 512); x = conv("conv15", x, 512); x = conv("conv16", x, 512)
    x = up("unpool2", x, 512); x = conv("conv17", x, 512); x = conv("conv18", x, 512); x = conv("conv19", x, 256)
    x = up("unpool3", x, 256); x = conv("conv20", x, 256); x = conv("conv21", x, 256); x = conv("conv22", x, 128)
    x = up("unpool4", x, 128); x = conv("conv23", x, 128); x = conv("conv24", x, 64)
    x = up("unpool5", x, 64); x = conv("conv25", x, 64)
    S.append(("conv26", "conv_bn_f32", [x], 3, num_classes))
    return S


def _spec(model, num_classes):
    return {"unet": unet_spec, "segnet": segnet_spec}[model](num_classes)


def unet_variable_shapes(cin=3, num_classes=2, model="unet"):
    ch = {"input": cin}
    shapes = OrderedDict()
    bn = 0
    for name, kind, inputs, k, co in _spec(model, num_classes):
        if kind == "pool":
            ch[name] = ch[inputs[0]]
        elif kind == "concat":
            ch[name] = ch[inputs[0]] + ch[inputs[1]]
        elif kind == "deconv":
            shapes[f"{name}/weights"] = (k, k, co, ch[inputs[0]])          # utils.py:264
            ch[name] = co
        else:
            shapes[f"{name}/weights"] = (k, k, ch[inputs[0]], co)          # utils.py:178, no bias (:180)
            if kind.startswith("conv_bn"):
                scope = "batch_normalization" if bn == 0 else f"batch_normalization_{bn}"
                bn += 1
                shapes[f"{scope}/gamma"] = (co,)
                shapes[f"{scope}/beta"] = (co,)
            ch[name] = co
    return shapes


def unet_init(cin=3, num_classes=2, seed=1234, init="ref", model="unet"):
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in unet_variable_shapes(cin, num_classes, model).items():
        if name.endswith("weights"):
            z = rng.standard_normal(shape, dtype=np.float32)
            if init == "ref":
                std = 0.01                                                   # utils.py:179
            else:
                fan_in = shape[0] * shape[1] * shape[2] if not name.startswith("unpool") else 4 * shape[3]
                std = float(np.sqrt(2.0 / fan_in))
            out[name] = (z * np.float32(std)).astype(np.float32)
        elif name.endswith("gamma"):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


class UNetOracle:
    def __init__(self, variables, num_classes=2, bf16_storage=False, bf16_grads=False, model="unet"):
        self.ncls = num_classes
        self.model = model
        self.bf16, self.bf16_grads = bf16_storage, bf16_grads
        self.vars = OrderedDict((k, torch.tensor(v, dtype=torch.float32, requires_grad=True)) for k, v in variables.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
        self.t = 0
        self.acts = OrderedDict()

    def _q(self, x):
        if not self.bf16:
            return x
        if self.bf16_grads:
            return _RoundBoth.apply(x)
        return x + (T.to_bf16_grid(x.detach()) - x.detach())

    def forward(self, x_u8):
        x = torch.as_tensor(np.asarray(x_u8), dtype=torch.float32)
        A = self.acts
        A.clear()
        A["input"] = x
        bn = 0
        for name, kind, inputs, k, co in _spec(self.model, self.ncls):
            if kind == "pool":
                A[name] = T.max_pool_2x2(A[inputs[0]])
            elif kind == "concat":
                A[name] = torch.cat([A[inputs[0]], A[inputs[1]]], dim=3)                 # utils.py:332
            elif kind == "deconv":
                src = A[inputs[0]]
                w = self.vars[f"{name}/weights"]
                if self.bf16:
                    w = w + (T.to_bf16_grid(w.detach()) - w.detach())
                y = T.conv2d_transpose_same(src, w, (src.shape[1] * 2, src.shape[2] * 2), 2)   # utils.py:275
                A[name] = self._q(y)
            elif kind.startswith("conv_bn"):
                scope = "batch_normalization" if bn == 0 else f"batch_normalization_{bn}"
                bn += 1
                g, b = self.vars[f"{scope}/gamma"], self.vars[f"{scope}/beta"]
                w = self.vars[f"{name}/weights"]
                # the CUDA path folds gamma / sqrt(1 + eps) into the bf16 weights: mirror that rounding point
                weff = w * (g / math.sqrt(1.0 + BN_EPS))
                if self.bf16:
                    weff = weff + (T.to_bf16_grid(weff.detach()) - weff.detach())
                y = T.conv2d_same(A[inputs[0]], weff) + b                                  # BN(conv(x)) (utils.py:196-201)
                if kind == "conv_bn_relu":
                    A[name] = self._q(T.relu(y))
                elif kind == "conv_bn":
                    A[name] = self._q(y)                                                    # relu=False (utils.py:194)
                else:
                    A[name] = y                                                             # fp32 logits (SegNet.py:80-81)
            else:
                w = self.vars[f"{name}/weights"]
                A[name] = T.conv2d_same(A[inputs[0]], w)                                   # fp32 logits
        logits = A[_spec(self.model, self.ncls)[-1][0]]
        return T.argmax_last(logits).unsqueeze(3), logits

    def loss(self, logits, labels_u8):
        lab = torch.as_tensor(np.asarray(labels_u8), dtype=torch.int64)
        onehot = torch.nn.functional.one_hot(lab, self.ncls).to(torch.float32)
        return T.softmax_cross_entropy_with_logits(logits, onehot).mean()

    def loss_and_grads(self, x, labels):
        for v in self.vars.values():
            v.grad = None
        _, logits = self.forward(x)
        loss = self.loss(logits, labels)
        loss.backward()
        return float(loss.detach()), logits.detach(), OrderedDict((k, v.grad.detach().clone()) for k, v in self.vars.items())

    def train_step(self, x, labels, lr=1e-4):
        loss, logits, grads = self.loss_and_grads(x, labels)
        self.t += 1
        with torch.no_grad():
            for k, p in self.vars.items():
                T.adam_tf_step(p, self.m[k], self.v[k], grads[k], self.t, lr=lr)
        return loss, logits, grads
