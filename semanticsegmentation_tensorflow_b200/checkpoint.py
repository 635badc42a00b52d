"""Checkpoint interchange: weights and optimizer slots under the reference's variable
names and layouts, as a NumPy `.npz` (the TF bundle format itself is out of scope — there is no
TensorFlow here to validate a writer against).

Names follow what `tf.train.Saver()` sees in the reference (`FCN.py:370-378`):
    <scope>/weights   HWIO fp32 (transposed conv: [kh,kw,Cout,Cin], FCN.py:143)   <scope>/biases   conv_t3/bias
    <var>/Adam, <var>/Adam_1       Adam first / second moment slots      <var>/Momentum   Momentum accumulator
    beta1_power, beta2_power        TF's non-slot accumulators (= beta^(t+1) after t applied steps), float32 as TF
                                    stores them; written for name parity only and never inverted (0.9^t underflows
                                    float32 after ~986 steps)
    global_step                     int64: the number of applied optimizer steps; `opt.t` is restored from this
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def _slots_sharded(net) -> bool:
    """True while the fused data-parallel exchange (dp.SymmetricAllReduce) keeps Adam's m / v sharded over
    the ranks and they have not been gathered since the last step."""
    ex = getattr(net, "exchange", None)
    return bool(ex is not None and getattr(ex, "fused", False) and getattr(ex, "slots_stale", False))


def state_dict(net, opt=None, train_step=None) -> "OrderedDict[str, np.ndarray]":
    """Optimizer slots sharded by the fused exchange are gathered first (a collective: call on every rank);
    without a `train_step` to gather through, exporting stale shards raises instead of writing zeros."""
    V = net.vars
    if opt is not None and V.m is not None and _slots_sharded(net):
        if train_step is not None:
            train_step.sync_optimizer_state()
        else:
            ex = net.exchange
            ex.gather_optimizer_state()
    out = OrderedDict()
    for name in V.slots:
        out[name] = V.view(V.p, name).detach().cpu().numpy().copy()
    if opt is not None and V.m is not None:
        momentum = not hasattr(opt, "beta1")
        for name in V.slots:
            out[f"{name}/Momentum" if momentum else f"{name}/Adam"] = V.view(V.m, name).detach().cpu().numpy().copy()
            if V.v is not None and not momentum:
                out[f"{name}/Adam_1"] = V.view(V.v, name).detach().cpu().numpy().copy()
        if not momentum:
            out["beta1_power"] = np.float32(np.float64(opt.beta1) ** (opt.t + 1))
            out["beta2_power"] = np.float32(np.float64(opt.beta2) ** (opt.t + 1))
        out["global_step"] = np.int64(opt.t)
    return out


def save_checkpoint(path, net, opt=None, train_step=None):
    """Data parallel with the fused exchange: the optimizer slots are sharded over the ranks and are gathered
    first -- a collective, so call this on every rank (only rank 0 needs to keep the file)."""
    np.savez(path, **state_dict(net, opt, train_step))


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
        slot = f"{name}/Adam" if f"{name}/Adam" in state else f"{name}/Momentum"
        if opt is not None and slot in state:
            if V.m is None:
                V.m = torch.zeros_like(V.p)
                if hasattr(opt, "beta1"):
                    V.v = torch.zeros_like(V.p)
            V.view(V.m, name).copy_(torch.as_tensor(np.asarray(state[slot], np.float32)))
            if f"{name}/Adam_1" in state and V.v is not None:
                V.view(V.v, name).copy_(torch.as_tensor(np.asarray(state[f"{name}/Adam_1"], np.float32)))
    if opt is not None:
        if "global_step" in state:
            opt.t = int(state["global_step"])
        elif "beta1_power" in state and hasattr(opt, "beta1"):
            # checkpoints written before `global_step` existed: invert beta1^(t+1) while float32 still resolves it
            b1p = float(state["beta1_power"])
            if not (b1p > 1e-30):
                raise ValueError("checkpoint has no global_step and beta1_power has underflowed: cannot recover the step count")
            opt.t = int(round(math.log(b1p) / math.log(opt.beta1))) - 1
    V.repack(net.ops)          # refresh the bf16 kernel-layout shadows


def load_checkpoint(path, net, opt=None, strict=True):
    with np.load(path) as z:
        load_state_dict(net, {k: z[k] for k in z.files}, opt, strict)
