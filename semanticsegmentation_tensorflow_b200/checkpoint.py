"""Checkpoint interchange: weights and optimizer slots under the reference's variable
names and layouts, as a NumPy `.npz` (the TF bundle format itself is out of scope — there is no
TensorFlow here to validate a writer against).

Names follow what `tf.train.Saver()` sees in the reference (`FCN.py:370-378`):
    <scope>/weights   HWIO fp32 (transposed conv: [kh,kw,Cout,Cin], FCN.py:143)   <scope>/biases   conv_t3/bias
    <var>/Adam, <var>/Adam_1       Adam first / second moment slots
    beta1_power, beta2_power        TF's non-slot accumulators (= beta^(t+1) after t applied steps)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def state_dict(net, opt=None) -> "OrderedDict[str, np.ndarray]":
    V = net.vars
    out = OrderedDict()
    for name in V.slots:
        out[name] = V.view(V.p, name).detach().cpu().numpy().copy()
    if opt is not None and V.m is not None:
        for name in V.slots:
            out[f"{name}/Adam"] = V.view(V.m, name).detach().cpu().numpy().copy()
            if V.v is not None:
                out[f"{name}/Adam_1"] = V.view(V.v, name).detach().cpu().numpy().copy()
        if hasattr(opt, "beta1"):
            out["beta1_power"] = np.float32(opt.beta1 ** (opt.t + 1))
            out["beta2_power"] = np.float32(opt.beta2 ** (opt.t + 1))
    return out


def save_checkpoint(path, net, opt=None, train_step=None):
    """`train_step`: pass the data-parallel TrainStep so that optimizer slots sharded over the ranks by the
    fused exchange (dp.SymmetricAllReduce) are gathered first -- a collective: call on every rank."""
    if train_step is not None:
        train_step.sync_optimizer_state()
    np.savez(path, **state_dict(net, opt))


def load_state_dict(net, state, opt=None, strict=True):
    V = net.vars
    missing = [n for n in V.slots if n not in state]
    if strict and missing:
        raise KeyError(f"checkpoint lacks variables: {missing[:5]}{'...' if len(missing) > 5 else ''}")
    for name in V.slots:
        if name not in state:
            continue
        arr = np.asarray(state[name], dtype=np.float32)
        if tuple(arr.shape) != tuple(V.slots[name].shape):
            raise ValueError(f"{name}: checkpoint shape {arr.shape} != {V.slots[name].shape}")
        V.view(V.p, name).copy_(torch.as_tensor(arr))
        if opt is not None and f"{name}/Adam" in state:
            if V.m is None:
                V.m = torch.zeros_like(V.p)
                V.v = torch.zeros_like(V.p)
            V.view(V.m, name).copy_(torch.as_tensor(np.asarray(state[f"{name}/Adam"], np.float32)))
            if f"{name}/Adam_1" in state:
                V.view(V.v, name).copy_(torch.as_tensor(np.asarray(state[f"{name}/Adam_1"], np.float32)))
    if opt is not None and "beta1_power" in state and hasattr(opt, "beta1"):
        opt.t = int(round(math.log(float(state["beta1_power"])) / math.log(opt.beta1))) - 1
    V.repack(net.ops)          # refresh the bf16 kernel-layout shadows


def load_checkpoint(path, net, opt=None, strict=True):
    with np.load(path) as z:
        load_state_dict(net, {k: z[k] for k in z.files}, opt, strict)
