"""CPU fp32 oracle for FCDenseNet (`Network/model/FCDenseNet.py:23-163`) built on the reference's bias-free shared
helpers (`Network/utils/utils.py:164-333`).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

PARITY UNPINNED (no reference tests / fixtures; TensorFlow not installable).

Written as the reference's own functions, one torch-CPU call per TF op:
    Conv2D_Block       tf.nn.conv2d SAME, no bias            utils.py:164-183
    Deconv2D_Block     tf.nn.conv2d_transpose 4x4 s2 SAME    utils.py:255-276
    Batch_Normalization  tf.layers.batch_normalization(x) with training=False and never-updated moving stats:
                       y = gamma * x / sqrt(1 + 1e-3) + beta  utils.py:300-301
    ReLU / Avg_Pooling (2x2 s2 VALID mean) / Dropout / Concat  utils.py:303,309,318,332
Variables are created in call order: `<name>/weights`, `batch_normalization[_k]/{gamma,beta}`."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import tf_ops as T
from .fcn_oracle import _RoundBoth

BN_EPS = 1e-3


class _Scope:
    """Creation-order variable bookkeeping (tf.get_variable / tf.layers naming)."""

    def __init__(self, variables=None):
        self.shapes = OrderedDict()
        self.vars = variables
        self.bn = 0

    def weights(self, name, shape):
        self.shapes[f"{name}/weights"] = tuple(shape)
        return None if self.vars is None else self.vars[f"{name}/weights"]

    def bn_params(self, c):
        scope = "batch_normalization" if self.bn == 0 else f"batch_normalization_{self.bn}"
        self.bn += 1
        self.shapes[f"{scope}/gamma"] = (c,)
        self.shapes[f"{scope}/beta"] = (c,)
        if self.vars is None:
            return None, None
        return self.vars[f"{scope}/gamma"], self.vars[f"{scope}/beta"]


class FCDenseNetOracle:
    def __init__(self, variables=None, num_classes=2, n_layers_per_blocks=(4, 5, 7, 10, 12, 15), growth_rate=16,
                 n_filters_first_conv=48, theta=0.5, cin=3, bf16_storage=False, bf16_grads=False):
        self.ncls, self.layers, self.growth, self.first, self.theta, self.cin = (num_classes, tuple(n_layers_per_blocks),
                                                                                 growth_rate, n_filters_first_conv, theta, cin)
        self.bf16, self.bf16_grads = bf16_storage, bf16_grads
        self.vars = None
        if variables is not None:
            self.vars = OrderedDict((k, torch.tensor(v, dtype=torch.float32, requires_grad=True)) for k, v in variables.items())
            self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
            self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.vars.items())
        self.t = 0
        self.acts = OrderedDict()

    # ---- storage rounding points of the CUDA path (bf16 tensors between kernels) ------------------
    def _q(self, x):
        if not self.bf16:
            return x
        if self.bf16_grads:
            return _RoundBoth.apply(x)
        return x + (T.to_bf16_grid(x.detach()) - x.detach())

    def _qw(self, w):
        return w if not self.bf16 else w + (T.to_bf16_grid(w.detach()) - w.detach())

    # ---- helpers (utils.py) ----------------------------------------------------------------------
    def _conv(self, S, x, cout, k, name, final=False):
        w = S.weights(name, (k, k, x.shape[3] if self.vars is not None else x, cout))
        if self.vars is None:
            return cout
        y = T.conv2d_same(x, self._qw(w))
        y = y if final else self._q(y)
        self.acts[name] = y
        return y

    def _deconv(self, S, x, cout, name):
        w = S.weights(name, (4, 4, cout, x.shape[3] if self.vars is not None else x))
        if self.vars is None:
            return cout
        y = self._q(T.conv2d_transpose_same(x, self._qw(w), (x.shape[1] * 2, x.shape[2] * 2), 2))
        self.acts[name] = y
        return y

    def _bn_relu(self, S, x):
        g, b = S.bn_params(x.shape[3] if self.vars is not None else x)
        if self.vars is None:
            return x
        return self._q(torch.relu(x * (g / math.sqrt(1.0 + BN_EPS)) + b))

    def _dropout(self, x, keep_prob, masks, key):
        if self.vars is None or keep_prob >= 1.0:
            return x
        return self._q(T.dropout(x, keep_prob, masks[key]))

    def _avg_pool(self, x):
        if self.vars is None:
            return x
        return self._q(F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous())

    def _cat(self, xs):
        if self.vars is None:
            return sum(xs)
        return torch.cat(xs, dim=3)

    # ---- FCDenseNet.py:23-163 -----------------------------------------------------------------------
    def _bottleneck(self, S, x, keep, masks, name):
        y = self._bn_relu(S, x)
        y = self._conv(S, y, 4 * self.growth, 1, name + "_conv1")
        y = self._dropout(y, keep, masks, name + "_conv1")
        y = self._bn_relu(S, y)
        y = self._conv(S, y, self.growth, 3, name + "_conv2")
        return self._dropout(y, keep, masks, name + "_conv2")

    def _transition(self, S, x, name):
        c = x.shape[3] if self.vars is not None else x
        y = self._bn_relu(S, x)
        y = self._conv(S, y, int(c * self.theta), 1, name + "_conv")
        return self._avg_pool(y)

    def _dense_block(self, S, x, n, keep, masks, name):
        layers = [x]
        layers.append(self._bottleneck(S, x, keep, masks, name + "bottleneck_layer_0"))
        for i in range(n):
            c = self._cat(layers)
            layers.append(self._bottleneck(S, c, keep, masks, name + "bottleneck_layer_" + str(i + 1)))
        return self._cat(layers)

    def _graph(self, S, x, keep, masks):
        x = self._conv(S, x, self.first, 3, "dense_init")
        blocks = []
        nb = len(self.layers)
        for b in range(nb):
            db = self._dense_block(S, x, self.layers[b], keep, masks, f"denseblock{b + 1}")
            blocks.append(db)
            if self.vars is not None:
                self.acts[f"denseblock{b + 1}"] = db
            if b < nb - 1:
                x = self._transition(S, db, f"transition_layer{b + 1}")
        x = blocks[-1]
        for u in range(nb - 1):
            skip = blocks[nb - 2 - u]
            up = self._deconv(S, x, skip.shape[3] if self.vars is not None else skip, f"transition_up{u + 1}")
            x = self._cat([up, skip])
        return self._conv(S, x, self.ncls, 1, "final_conv", final=True)

    def variable_shapes(self):
        vars_, self.vars = self.vars, None
        S = _Scope(None)
        self._graph(S, self.cin, 1.0, None)
        self.vars = vars_
        return S.shapes

    def forward(self, x_u8, keep_prob=1.0, masks=None):
        x = torch.as_tensor(np.asarray(x_u8), dtype=torch.float32)
        self.acts.clear()
        logits = self._graph(_Scope(self.vars), x, keep_prob, masks or {})
        return T.argmax_last(logits).unsqueeze(3), logits

    def loss(self, logits, labels_u8):
        lab = torch.as_tensor(np.asarray(labels_u8), dtype=torch.int64)
        onehot = F.one_hot(lab, self.ncls).to(torch.float32)
        return T.softmax_cross_entropy_with_logits(logits, onehot).mean()

    def loss_and_grads(self, x, labels, keep_prob=1.0, masks=None):
        for v in self.vars.values():
            v.grad = None
        _, logits = self.forward(x, keep_prob, masks)
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


def densenet_init(shapes, seed=1234, init="ref"):
    """weights N(0, 0.01^2) (utils.py:179,266), gamma 1, beta 0; 'he' = std sqrt(2 / fan_in) for visibility tests."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("weights"):
            z = rng.standard_normal(shape, dtype=np.float32)
            if init == "ref":
                std = 0.01
            else:
                fan_in = 4 * shape[3] if name.startswith("transition_up") else shape[0] * shape[1] * shape[2]
                std = float(np.sqrt(2.0 / fan_in))
            out[name] = (z * np.float32(std)).astype(np.float32)
        elif name.endswith("gamma"):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out
