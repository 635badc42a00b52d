"""FCDenseNet ("Tiramisu") of the reference (`Network/model/FCDenseNet.py:23-163`) on the B200 kernels, behind the
reference's builder call shape `FCDenseNet(x, keep_prob, num_classes) -> (pred, logits)` (via `.create()`).

    bottleneck_layer   BN -> ReLU -> 1x1 conv (4*growth) -> Dropout -> BN -> ReLU -> 3x3 conv (growth) -> Dropout   :23-35
    Transition_Layer   BN -> ReLU -> 1x1 conv (theta * C) -> Avg_Pooling 2x2                                       :37-46
    DenseBlock         x, then n+1 bottlenecks each fed by Concat(everything so far); output = Concat(all)         :48-60
    decoder            Deconv2D_Block(4x4 s2, to the skip's channel count) -> Concat([up, dense_block_k]) x 5       :141-154
    head               1x1 Conv2D_Block to num_classes (no bias), argmax                                             :157-163

The helpers are the bias-free ones of `Network/utils/utils.py` (Conv2D_Layer :164-183, Deconv2D_Layer :255-276,
Batch_Normalization = inference-mode affine :300-301, Avg_Pooling :309, Dropout :318, Concat :332).

How it maps onto the tcgen05 kernels ("FCDenseNet"):
  * Channel counts here are 48 + 16 j, 140, 174, 348, 430 ... -- not multiples of 64, some not of 8.  Every tensor is
    stored PHYSICALLY with each channel segment padded to a multiple of 8 and the total to a multiple of 64; pad
    channels are always zero.  Weights are remapped logical -> physical (zero rows / columns at the pads) before
    the usual bf16 packing, gradients back; so each conv is the ordinary 64-channel-chunk GEMM.
  * Concat(layers_concat) of a dense block is zero-copy: one buffer per block ("root"), every earlier Concat is a
    channel PREFIX of it, and the pre-activation BN + ReLU kernel reads that prefix in place (row pitch = buffer
    width).  A produced tensor is copied once into its slot.
  * Only the concat gradient accumulates: BN/ReLU backward adds into the root's gradient prefix.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import plan as P
from .fcn import _Feeds
from .graph import BN_SCALE, _Vars
from .ops import Ops, _p, _stream, conv_flops, deconv_flops
from .overlap import SideStream


@dataclass
class DNode:
    name: str
    kind: str                     # conv | deconv | bnrelu | dropout | avgpool | concat
    inputs: List[str] = field(default_factory=list)
    k: int = 1
    cout: int = 0
    stride: int = 1
    relu: bool = False            # bnrelu: ReLU after the affine
    keep: float = 1.0             # dropout keep_prob (None = the net's keep_prob feed)
    bn_scope: str = ""


class DenseBuilder:
    """Collects calls with the reference's helper names (utils.py) into a node list."""

    def __init__(self):
        self.nodes: List[DNode] = []
        self._bn = 0
        self._auto = 0

    def _name(self, prefix):
        self._auto += 1
        return f"{prefix}_{self._auto}"

    def Conv2D_Block(self, x, num_filters, filter_height=3, filter_width=3, stride=1, name=None):
        assert filter_height == filter_width and stride == 1
        self.nodes.append(DNode(name, "conv", [x], k=filter_height, cout=num_filters))
        return name

    def Batch_Normalization(self, x):
        scope = "batch_normalization" if self._bn == 0 else f"batch_normalization_{self._bn}"
        self._bn += 1
        name = self._name("bn")
        self.nodes.append(DNode(name, "bnrelu", [x], bn_scope=scope))
        return name

    def ReLU(self, x):
        n = self.nodes[-1]
        assert n.name == x and n.kind == "bnrelu" and not n.relu, "ReLU is built fused behind Batch_Normalization"
        n.relu = True
        return x

    def Dropout(self, x, keep_prob):
        name = self._name("dropout")
        self.nodes.append(DNode(name, "dropout", [x], keep=keep_prob))
        return name

    def Avg_Pooling(self, x, name):
        self.nodes.append(DNode(name, "avgpool", [x]))
        return name

    def Concat(self, xs, axis=-1, name=None):
        self.nodes.append(DNode(name, "concat", list(xs)))
        return name

    def Deconv2D_Block(self, x, out_channels, name=None):
        self.nodes.append(DNode(name, "deconv", [x], k=4, cout=out_channels, stride=2))
        return name


def fcdensenet_nodes(num_classes=2, keep_prob=None, n_layers_per_blocks=(4, 5, 7, 10, 12, 15), growth_rate=16,
                     n_filters_first_conv=48, theta=0.5, cin=3):
    """The graph of FCDenseNet.py:83-163, call for call.  Returns (nodes, channels {tensor: C})."""
    g = DenseBuilder()
    ch = {"input": cin}

    def track(name, c):
        ch[name] = c
        return name

    def bottleneck_layer(x, name):                                              # FCDenseNet.py:23-35
        y = g.ReLU(track(g.Batch_Normalization(x), ch[x]))
        y = track(g.Conv2D_Block(y, 4 * growth_rate, 1, 1, name=name + "_conv1"), 4 * growth_rate)
        y = track(g.Dropout(y, keep_prob), ch[y])
        y = g.ReLU(track(g.Batch_Normalization(y), ch[y]))
        y = track(g.Conv2D_Block(y, growth_rate, name=name + "_conv2"), growth_rate)
        return track(g.Dropout(y, keep_prob), growth_rate)

    def Transition_Layer(x, name):                                              # :37-46
        y = g.ReLU(track(g.Batch_Normalization(x), ch[x]))
        y = track(g.Conv2D_Block(y, int(ch[x] * theta), 1, 1, name=name + "_conv"), int(ch[x] * theta))
        return track(g.Avg_Pooling(y, name=name + "avg_pool"), ch[y])

    def DenseBlock(x, n, name):                                                 # :48-60
        layers = [x]
        y = bottleneck_layer(x, name + "bottleneck_layer_0")
        layers.append(y)
        for i in range(n):
            c = track(g.Concat(layers, name=name + "bottleneck_layer_concatenate_" + str(i + 1)), sum(ch[t] for t in layers))
            y = bottleneck_layer(c, name + "bottleneck_layer_" + str(i + 1))
            layers.append(y)
        return track(g.Concat(layers, name=name + "bottleneck_layer_concatenate_final"), sum(ch[t] for t in layers))

    x = track(g.Conv2D_Block("input", n_filters_first_conv, name="dense_init"), n_filters_first_conv)     # :91
    blocks = []
    nb = len(n_layers_per_blocks)
    for b in range(nb):                                                         # :95-129
        db = DenseBlock(x, n_layers_per_blocks[b], f"denseblock{b + 1}")
        blocks.append(db)
        if b < nb - 1:
            x = Transition_Layer(db, f"transition_layer{b + 1}")
    x = blocks[-1]
    for u in range(nb - 1):                                                     # :141-154
        skip = blocks[nb - 2 - u]
        up = track(g.Deconv2D_Block(x, ch[skip], name=f"transition_up{u + 1}"), ch[skip])
        x = track(g.Concat([up, skip], name=f"tu_db_concat{u + 1}"), ch[up] + ch[skip])
    track(g.Conv2D_Block(x, num_classes, 1, 1, name="final_conv"), num_classes)                            # :157
    return g.nodes, ch


def densenet_variable_shapes(nodes, ch):
    """Ordered {name: shape} in creation order: `<scope>/weights` (HWIO; deconv [k,k,Cout,Cin], utils.py:264), and
    `batch_normalization_k/{gamma,beta}` over the BN input's channels."""
    shapes = OrderedDict()
    for n in nodes:
        ci = ch[n.inputs[0]] if n.inputs else 0
        if n.kind == "conv":
            shapes[f"{n.name}/weights"] = (n.k, n.k, ci, n.cout)
        elif n.kind == "deconv":
            shapes[f"{n.name}/weights"] = (n.k, n.k, n.cout, ci)
        elif n.kind == "bnrelu":
            shapes[f"{n.bn_scope}/gamma"] = (ci,)
            shapes[f"{n.bn_scope}/beta"] = (ci,)
    return shapes


def densenet_init(shapes, seed=1234, init="ref"):
    """weights N(0, 0.01^2) (utils.py:179,266), gamma 1, beta 0; 'he': std sqrt(2 / fan_in) for visibility tests."""
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


def _r8(c):
    return -(-c // 8) * 8


def _r64(c):
    return -(-c // 64) * 64


class _Layout:
    """Physical channel layout of a tensor: logical segments, each padded to a multiple of 8."""

    def __init__(self, segs, align=64):
        self.segs = list(segs)
        self.logical = sum(self.segs)
        self.used = sum(_r8(s) for s in self.segs)           # physical channels carrying data (+ inner pads)
        self.cp = -(-self.used // align) * align
        cmap = np.full(self.cp, -1, np.int32)
        lo = po = 0
        for s in self.segs:
            cmap[po:po + s] = np.arange(lo, lo + s, dtype=np.int32)
            lo += s
            po += _r8(s)
        self.cmap = cmap


class FCDenseNet(_Feeds):
    """`FCDenseNet(x, keep_prob, num_classes)` (FCDenseNet.py:83) with the FCN-compatible training interface."""

    def __init__(self, x, keep_prob=1.0, num_classes=2, variables=None, init="ref", seed=1234, world_size=1,
                 n_layers_per_blocks=(4, 5, 7, 10, 12, 15), growth_rate=16, n_filters_first_conv=48, theta=0.5,
                 dropout_seed=42):
        if not torch.cuda.is_available():
            raise RuntimeError("FCDenseNet needs a CUDA (sm_100a) device: the segmentation ops have no CPU fallback")
        x = torch.as_tensor(x)
        self.device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.ops = Ops(self.device)
        self.x = self._as_image(x).to(self.device)
        self.num_classes = int(num_classes)
        if not 2 <= self.num_classes <= 64:
            raise ValueError(f"num_classes must be in [2, 64] (got {self.num_classes})")
        self.N, self.H, self.W, self.Cin = self.x.shape
        down = 2 ** (len(n_layers_per_blocks) - 1)
        if self.H % down or self.W % down:
            raise ValueError(f"image size {self.H}x{self.W} must be a multiple of {down}")
        self.keep_prob = float(keep_prob)
        self.world_size = world_size
        self.dropout_seed = dropout_seed
        self.step_count = 0
        self.injected_masks = None          # {dropout node name: u8 mask tensor [N,H,W,Cp]} for parity runs
        self._init_feeds()
        self.nodes, self.ch = fcdensenet_nodes(self.num_classes, None, n_layers_per_blocks, growth_rate,
                                               n_filters_first_conv, theta, self.Cin)
        self.by_name = {n.name: n for n in self.nodes}
        shapes = densenet_variable_shapes(self.nodes, self.ch)
        values = variables if variables is not None else densenet_init(shapes, seed, init)
        self.vars = _Vars(shapes, self.device, values)
        self.vars._repack = self._repack
        self._plan()
        self._repack(self.ops)
        self._ran_forward = False
        self.side = SideStream(self.device, enabled=False)      # (single stream: no side-stream overlap yet)
        self.wside = SideStream(self.device, enabled=False)

    # ---- planning: layouts, concat roots, buffers --------------------------------------------------
    def _plan(self):
        dev, N, bf = self.device, self.N, torch.bfloat16
        nodes = self.nodes
        # concat aliasing: a Concat whose inputs are a prefix of a later Concat's inputs is a channel prefix of it
        root = {}
        concats = [n for n in nodes if n.kind == "concat"]
        for i, n in enumerate(concats):
            r = n
            for m in concats[i + 1:]:
                if m.inputs[:len(r.inputs)] == r.inputs:
                    r = m
            root[n.name] = r.name
        self.root = root
        # spatial size and layout of every tensor
        hw = {"input": (self.H, self.W)}
        lay = {"input": _Layout([self.Cin])}
        for n in nodes:
            h, w = hw[n.inputs[0]]
            if n.kind == "avgpool":
                h, w = h // 2, w // 2
            elif n.kind == "deconv":
                h, w = h * 2, w * 2
            hw[n.name] = (h, w)
            if n.kind == "concat":
                # A concat that only feeds a transposed conv (the decoder's Concat([up, skip]), FCDenseNet.py:146-147) is that
                # layer's GEMM-N in dgrad / wgrad: 320 / 448 / 576 / 704 channels would mean 64-wide tiles (bound by shared-
                # memory operand reads, measured 0.28 PFLOP/s); padded to a multiple of 128 they take 128 / 256-wide tiles
                readers = [m for m in nodes if n.name in m.inputs]
                wide = bool(readers) and all(m.kind == "deconv" for m in readers)
                lay[n.name] = _Layout([s for t in n.inputs for s in lay[t].segs], align=128 if wide else 64)
            elif n.kind in ("conv", "deconv"):
                lay[n.name] = _Layout([n.cout])
            else:
                lay[n.name] = _Layout(lay[n.inputs[0]].segs)
        self.hw, self.lay = hw, lay
        # root buffers and member slots.  member[producer] = (root name, physical offset); a Dropout between the
        # producer and the Concat works in place on the producer's tensor, so the slot is filled after the LAST
        # node of that chain (copy_after)
        self.buf, self.gbufs, self.member, self.copy_after = {}, {}, {}, {}
        for n in concats:
            if root[n.name] != n.name:
                continue
            h, w = hw[n.name]
            self.buf[n.name] = torch.zeros((N, h, w, lay[n.name].cp), dtype=bf, device=dev)
            self.gbufs[n.name] = torch.zeros((N, h, w, lay[n.name].cp), dtype=bf, device=dev)
            off = 0
            for t in n.inputs:
                b = self._base(t)
                assert b not in self.member, f"{b} is a member of two concat buffers"
                self.member[b] = (n.name, off)
                self.copy_after[t] = b
                off += lay[t].used
        # Dropout nodes applied on the fly by their only reader: the next BN/ReLU (bottleneck conv1 -> Dropout -> BN) or the
        # copy into the concat slot (conv2 -> Dropout -> Concat); the others run segk_dropout in place
        self.node_index = {n.name: i for i, n in enumerate(nodes)}
        self.drop_fused, self.bn_drop, self.copy_drop = {}, {}, {}
        for d in nodes:
            if d.kind != "dropout":
                continue
            b = self._base(d.name)
            users = [m for m in nodes if d.name in m.inputs]
            if self.by_name[b].kind not in ("conv", "deconv") or not users:
                continue
            if len(users) == 1 and users[0].kind == "bnrelu" and b not in self.member and d.inputs[0] == b:
                self.drop_fused[d.name] = "bn"
                self.bn_drop[users[0].name] = d
            elif all(m.kind == "concat" for m in users) and self.copy_after.get(d.name) == b and d.inputs[0] == b:
                self.drop_fused[d.name] = "copy"
                self.copy_drop[b] = d
        last = nodes[-1].name
        self.act, self.g = {}, {}
        for n in nodes:
            h, w = hw[n.name]
            if n.kind == "concat":
                continue
            if n.kind == "dropout":
                continue                                         # in place on its input
            if n.kind == "avgpool" and n.name in self.member:
                continue                                         # written straight into its concat slot
            if n.name == last:
                self.act[n.name] = torch.empty((N, h, w, n.cout), dtype=torch.float32, device=dev)
                self.g[n.name] = torch.empty((N, h, w, n.cout), dtype=torch.float32, device=dev)
                continue
            cp = lay[n.name].cp
            self.act[n.name] = torch.zeros((N, h, w, cp), dtype=bf, device=dev)
            self.g[n.name] = torch.zeros((N, h, w, cp), dtype=bf, device=dev)
        # physical parameter buffers
        V = self.vars
        self.weff, self.gw_phys, self.scale_p, self.shift_p, self.dscale_p, self.dshift_p, self.cmap_dev = {}, {}, {}, {}, {}, {}, {}
        gw_max = 0
        for n in nodes:
            if n.kind in ("conv", "deconv"):
                cin_p = lay[n.inputs[0]].cp if n.inputs[0] != "input" else self.Cin
                cout_p = n.cout if n.name == last else _r64(n.cout)
                shape = (n.k, n.k, cin_p, cout_p) if n.kind == "conv" else (n.k, n.k, cout_p, cin_p)
                self.weff[n.name] = torch.zeros(shape, dtype=torch.float32, device=dev)
                gw_max = max(gw_max, int(np.prod(shape)))
                if n.inputs[0] != "input":
                    self.cmap_dev[n.name] = torch.as_tensor(lay[n.inputs[0]].cmap).to(dev)
        # BN parameters of ALL layers live in four flat physical buffers; one global map per buffer points into the
        # variable arena (logical gamma / beta), so a step needs two gathers and two scatters, not four per layer
        bns = [n for n in nodes if n.kind == "bnrelu"]
        total_p = sum(lay[n.inputs[0]].cp for n in bns)
        flat = {k: torch.zeros(total_p, dtype=torch.float32, device=dev) for k in ("scale", "shift", "dscale", "dshift")}
        gmap, bmap = np.full(total_p, -1, np.int32), np.full(total_p, -1, np.int32)
        o = 0
        for n in bns:
            L = lay[n.inputs[0]]
            go, bo = V.slots[f"{n.bn_scope}/gamma"].offset, V.slots[f"{n.bn_scope}/beta"].offset
            m = L.cmap >= 0
            gmap[o:o + L.cp][m] = go + L.cmap[m]
            bmap[o:o + L.cp][m] = bo + L.cmap[m]
            self.scale_p[n.name], self.shift_p[n.name] = flat["scale"][o:o + L.cp], flat["shift"][o:o + L.cp]
            self.dscale_p[n.name], self.dshift_p[n.name] = flat["dscale"][o:o + L.cp], flat["dshift"][o:o + L.cp]
            o += L.cp
        self.bn_flat = flat
        self.bn_gmap, self.bn_bmap = torch.as_tensor(gmap).to(dev), torch.as_tensor(bmap).to(dev)
        self.gw_scratch = torch.empty(gw_max, dtype=torch.float32, device=dev)
        self.bn_ws = self.ops.bn_act_bwd_workspace(max(l.cp for l in lay.values()), dev)
        self.logits = self.act[last]
        self.dlogits = self.g[last]
        self.dlogits_bf16 = torch.empty(self.logits.shape, dtype=bf, device=dev)
        npix = N * self.H * self.W
        self.pred_u8 = torch.empty((N, self.H, self.W), dtype=torch.uint8, device=dev)
        self.loss_sum = torch.zeros(2, dtype=torch.float32, device=dev)
        self.xent_ws = self.ops.xent_workspace(npix, dev)
        self.labels = torch.zeros((N, self.H, self.W), dtype=torch.uint8, device=dev)
        self.route = {}
        for n in nodes:
            if n.kind == "conv":
                if n.inputs[0] == "input":
                    if not (n.k == 3 and self.Cin in (1, 3, 4) and n.cout <= 64):
                        raise ValueError(f"{n.name}: first layer must be 3x3 on 1/3/4 channels to <= 64 (no fallback)")
                    self.route[n.name] = "first"
                elif n.name == last:
                    if not (n.k == 1 and n.cout in (2, 4, 8)):
                        raise ValueError(f"{n.name}: head must be 1x1 to 2/4/8 classes (no fallback)")
                    self.route[n.name] = "small"
                else:
                    self.route[n.name] = "tc"

    # ---- views --------------------------------------------------------------------------------------
    def _base(self, name):
        """The producer behind a chain of (in-place) Dropout nodes."""
        n = self.by_name.get(name)
        while n is not None and n.kind == "dropout":
            name = n.inputs[0]
            n = self.by_name.get(name)
        return name

    def _storage(self, name, grad=False):
        """(tensor, physical channel offset, data channels) where tensor `name` (or its gradient) is READ: the slot
        of its concat buffer if it lives in one, a prefix of the root buffer for a Concat, else its own tensor."""
        if name == "input":
            return self.x, 0, self.Cin
        b = self._base(name)
        n = self.by_name[b]
        if n.kind == "concat":               # (a block buffer is read from itself even when a decoder concat holds a copy)
            return (self.gbufs if grad else self.buf)[self.root[b]], 0, self.lay[b].used
        if b in self.member:
            r, off = self.member[b]
            return (self.gbufs if grad else self.buf)[r], off, self.lay[b].used
        return (self.g if grad else self.act)[b], 0, self.lay[b].used

    def _dense_input(self, name):
        """A conv reads its input as a dense [N,H,W,Cp] tensor: a BN/ReLU output, a transposed-conv output or a whole
        concat buffer."""
        b = self._base(name)
        n = self.by_name[b]
        if n.kind == "concat":
            if self.root[b] != b:
                raise ValueError(f"{name}: a conv cannot read a concat prefix directly")
            return self.buf[b]
        if b in self.member:
            raise ValueError(f"{name}: a conv cannot read a concat member directly")
        return self.act[b]

    # ---- parameters: logical -> physical -> bf16 kernel layouts -----------------------------------------
    def _repack(self, ops, only=None):
        V = self.vars
        # scale = gamma / sqrt(1 + eps), shift = beta, for every BN layer at once (physical order, pads 0)
        ops.gather_f32(V.p, self.bn_gmap, self.bn_flat["scale"], mul=BN_SCALE)
        ops.gather_f32(V.p, self.bn_bmap, self.bn_flat["shift"])
        for n in self.nodes:
            if n.kind in ("conv", "deconv"):
                if only is not None and n.name not in only:
                    continue
                w = V.param(f"{n.name}/weights")
                weff = self.weff[n.name]
                cm = self.cmap_dev.get(n.name)
                if n.kind == "conv":          # [k,k,Cin,Cout]: rows = Cin (mapped), cols = Cout (zero-padded)
                    ops.remap_weights(w, weff, amap=cm, bmap=None, to_phys=True)
                    r = self.route[n.name]
                    if r == "tc":
                        V.wk[n.name], V.wd[n.name] = ops.pack_conv_weights(weff, V.wk.get(n.name), V.wd.get(n.name))
                    elif r == "first":
                        V.wk[n.name] = ops.pack_im2col_weights(weff, V.wk.get(n.name))
                else:                          # [k,k,Cout,Cin]: rows = Cout (padded), cols = Cin (mapped)
                    ops.remap_weights(w, weff, amap=None, bmap=cm, to_phys=True)
                    V.wk[n.name], V.wd[n.name] = ops.pack_deconv_weights(weff, n.stride, V.wk.get(n.name), V.wd.get(n.name))

    # ---- forward ------------------------------------------------------------------------------------
    def _dropout_args(self, n, idx):
        keep = self.keep_prob if n.keep is None else float(n.keep)
        mask = None if self.injected_masks is None else self.injected_masks.get(n.name)
        seed = (self.dropout_seed * 1000003 + self.step_count) * 1024 + idx
        return keep, seed, mask

    def _drop_of(self, d):
        """(keep_prob, seed, mask) of Dropout node d, or None when it is inactive (keep_prob 1)."""
        keep, seed, mask = self._dropout_args(d, self.node_index[d.name])
        return (keep, seed, mask) if keep < 1.0 else None

    def _to_root(self, name):
        """After node `name` ran: if it ends the chain feeding a concat slot, copy the produced tensor into the slot."""
        b = self.copy_after.get(name)
        if b is None:
            return
        r, off = self.member[b]
        n = self.by_name[b]
        if n.kind == "avgpool":
            return                               # written in place
        if n.kind == "concat":                   # a whole block buffer inside a decoder concat
            src, c = self.buf[self.root[b]], self.lay[b].used
        else:
            src, c = self.act[b], self.lay[b].used
        drop = None
        d = self.copy_drop.get(b)
        if d is not None and d.name == name:
            a = self._drop_of(d)
            drop = None if a is None else (1,) + a
        self.ops.channel_copy(src, 0, self.buf[r], off, c, drop=drop)

    def forward(self):
        ops, V = self.ops, self.vars
        for idx, n in enumerate(self.nodes):
            if n.kind == "concat":
                if self.root[n.name] == n.name:
                    pass                         # members were copied as they were produced
            elif n.kind == "bnrelu":
                src, off, c = self._storage(n.inputs[0])
                assert off == 0, "BN/ReLU reads a channel prefix"
                d = self.bn_drop.get(n.name)
                ops.bn_act_fwd(src, c, self.act[n.name], self.scale_p[n.name], self.shift_p[n.name], relu=n.relu,
                               drop=None if d is None else self._drop_of(d))
            elif n.kind == "dropout":
                keep, seed, mask = self._dropout_args(n, idx)
                if keep < 1.0 and n.name not in self.drop_fused:
                    t = self.act[self._base(n.name)]
                    ops.dropout(t, t, keep, seed, mask)
            elif n.kind == "avgpool":
                src, soff, c = self._storage(n.inputs[0])
                assert soff == 0
                if n.name in self.member:        # straight into the next block's buffer
                    r, off = self.member[n.name]
                    dst = self.buf[r]
                    h, w = self.hw[n.inputs[0]]
                    ops.call("segk_avgpool2x2_fwd", _p(src), src.shape[-1], dst.data_ptr() + 2 * off, dst.shape[-1], self.N, h, w,
                             c, _stream())
                else:
                    ops.avgpool_fwd(src, c, self.act[n.name])
            elif n.kind == "deconv":
                x = self._dense_input(n.inputs[0])
                ops.deconv2d_fwd(x, V.wk[n.name], None, self.act[n.name], n.k, n.stride)
            else:
                r = self.route[n.name]
                out = self.act[n.name]
                if r == "first":
                    ops.conv2d_first_fwd(self.x, V.wk[n.name], None, out, n.k, n.k, relu=False)
                elif r == "small":
                    ops.conv2d_small_fwd(self._dense_input(n.inputs[0]), self.weff[n.name], None, out, relu=False)
                else:
                    ops.conv2d_fwd(self._dense_input(n.inputs[0]), V.wk[n.name], None, out, n.k, n.k, relu=False)
            self._to_root(n.name)
        self._ran_forward = True
        return self.logits

    def create(self):
        self.forward()
        self.ops.softmax_infer(self.logits, None, None, self.pred_u8)       # tf.argmax (FCDenseNet.py:159)
        return self.pred_u8.to(torch.int64).unsqueeze(3), self.logits

    def loss(self, annotation=None, with_grad=False):
        if annotation is not None:
            self.labels = self._as_labels(annotation)
        if not self._ran_forward:
            self.forward()
        npix = self.N * self.H * self.W
        self.ops.softmax_xent(self.logits, self.labels, self.dlogits if with_grad else None, self.pred_u8,
                              self.loss_sum, None, self.xent_ws, 1.0 / (npix * self.world_size))
        return self.loss_sum[1]

    def confusion_matrix(self):
        cm = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.ops.confusion_matrix(self.labels, self.pred_u8, cm)
        return cm.view(2, 2)

    def infer(self, image=None):
        if image is not None:
            self.feed({self.image: image})
        kp, self.keep_prob = self.keep_prob, 1.0
        try:
            self.forward()
        finally:
            self.keep_prob = kp
        prob = torch.empty_like(self.logits)
        mask = torch.empty((self.N, self.H, self.W), dtype=torch.uint8, device=self.device)
        self.ops.softmax_infer(self.logits, prob, mask)
        self.mark_step_end()
        return prob, mask

    # ---- backward -----------------------------------------------------------------------------------
    def backward(self, after_layer=None):
        ops, V = self.ops, self.vars
        has = set()          # gradient storages (root buffers / own tensors) that already hold a contribution
        pulled = set()       # members whose gradient slot has been copied out of the root's gradient

        def grad_of(name):
            """Dense gradient tensor of a produced tensor.  One that lives in a concat buffer gets it from its slot of
            the root's gradient: copied out once (a block buffer inside a decoder concat: added into its own root)."""
            b = self._base(name)
            n = self.by_name[b]
            if b in self.member and b not in pulled:
                r, off = self.member[b]
                assert r in has, f"{b}: the concat buffer {r} has no gradient yet"
                pulled.add(b)
                if n.kind == "concat":
                    dst = self.gbufs[self.root[b]]
                    ops.channel_copy(self.gbufs[r], off, dst, 0, self.lay[b].used, accumulate=self.root[b] in has)
                    has.add(self.root[b])
                else:
                    d, drop = self.copy_drop.get(b), None
                    if d is not None:            # DropoutGrad of the Dropout between this conv and its concat slot
                        a = self._drop_of(d)
                        drop = None if a is None else (2,) + a
                    ops.channel_copy(self.gbufs[r], off, self.g[b], 0, self.lay[b].used, drop=drop)
            if n.kind == "concat":
                return self.gbufs[self.root[b]]
            return self.g[b]

        def into(name):
            """(gradient storage, data channels of that storage, accumulate?) for a consumer writing the gradient of the
            tensor it READ as `name` (see _storage)."""
            b = self._base(name)
            n = self.by_name[b]
            if n.kind == "concat":
                key = self.root[b]
                t = self.gbufs[key]
            elif b in self.member:
                key, off = self.member[b]
                assert off == 0, "only the leading member of a concat buffer is read on its own"
                t = self.gbufs[key]
            else:
                key, t = b, self.g[b]
            full = self.lay[key].used if key in self.gbufs else self.lay[b].used
            acc = key in has
            has.add(key)
            return t, full, acc

        last = self.nodes[-1].name
        has.add(last)
        done = []
        for idx in range(len(self.nodes) - 1, -1, -1):
            n = self.nodes[idx]
            if n.kind == "concat":
                if n.name in self.member:
                    grad_of(n.name)                # add the decoder concat's slice to this block buffer's gradient
                continue
            if n.kind == "dropout":
                keep, seed, mask = self._dropout_args(n, idx)
                if keep < 1.0 and n.name not in self.drop_fused:
                    g = grad_of(n.name)
                    ops.dropout(g, g, keep, seed, mask)
                continue
            src = n.inputs[0]
            if n.kind == "bnrelu":
                dy = grad_of(n.name)
                xs, _, c = self._storage(src)
                dx, full, acc = into(src)
                if not acc and c < full:
                    dx.zero_()                      # the first writer covers only a prefix: the rest starts at zero
                    acc = True
                d = self.bn_drop.get(n.name)
                drop = None if d is None else self._drop_of(d)
                assert drop is None or not acc, "a tensor behind a fused Dropout has one reader"
                ops.bn_act_bwd(dy, self.act[n.name], xs, dx, c, self.scale_p[n.name], self.dscale_p[n.name],
                               self.dshift_p[n.name], self.bn_ws, relu=n.relu, accumulate=acc, drop=drop)
                continue
            if n.kind == "avgpool":
                if n.name in self.member:
                    r, off = self.member[n.name]
                    assert r in has
                    pulled.add(n.name)
                    gsrc, lddy = self.gbufs[r].data_ptr() + 2 * off, self.gbufs[r].shape[-1]
                else:
                    gsrc, lddy = grad_of(n.name).data_ptr(), self.g[n.name].shape[-1]
                dx, _, acc = into(src)
                assert not acc
                h, w = self.hw[src]
                ops.call("segk_avgpool2x2_bwd", gsrc, lddy, _p(dx), dx.shape[-1], self.N, h, w, self.lay[src].used, _stream())
                continue
            # conv / deconv
            dz = grad_of(n.name) if n.name != last else self.dlogits
            gw = V.grad(f"{n.name}/weights")
            weff = self.weff[n.name]
            gwp = self.gw_scratch[:weff.numel()].view(weff.shape)
            cm = self.cmap_dev.get(n.name)
            if n.kind == "deconv":
                x = self._dense_input(src)
                ops.deconv2d_wgrad(x, dz, gwp, n.k, n.stride)
                ops.remap_weights(gw, gwp, amap=None, bmap=cm, to_phys=False)
                dx, _, acc = into(src)
                if acc:
                    raise NotImplementedError(f"{n.name}: transposed conv input with an earlier gradient contribution")
                ops.deconv2d_dgrad(dz, V.wd[n.name], dx, n.k, n.stride)
            else:
                r = self.route[n.name]
                if r == "first":
                    ops.conv2d_first_wgrad(self.x, dz, gwp, n.k, n.k)
                    ops.remap_weights(gw, gwp, amap=None, bmap=None, to_phys=False)
                elif r == "small":
                    x = self._dense_input(src)
                    dzb = ops.cast_to_bf16(dz, self.dlogits_bf16)
                    ops.conv2d_small_wgrad(x, dzb, gwp)
                    ops.remap_weights(gw, gwp, amap=cm, bmap=None, to_phys=False)
                    dx, _, acc = into(src)
                    if acc:
                        raise NotImplementedError(f"{n.name}: head input with an earlier gradient contribution")
                    ops.conv2d_small_dgrad(dzb, weff, dx)
                else:
                    x = self._dense_input(src)
                    ops.conv2d_wgrad(x, dz, gwp, n.k, n.k)
                    ops.remap_weights(gw, gwp, amap=cm, bmap=None, to_phys=False)
                    dx, _, acc = into(src)
                    ops.conv2d_dgrad(dz, V.wd[n.name], dx, n.k, n.k, residual=dx if acc else None)
            done.append(n.name)
        # d gamma = d scale / sqrt(1 + eps), d beta = d shift: all BN layers at once, into the gradient arena
        ops.scatter_f32(self.bn_flat["dscale"], self.bn_gmap, V.g, mul=BN_SCALE)
        ops.scatter_f32(self.bn_flat["dshift"], self.bn_bmap, V.g)
        if after_layer is not None:          # (data parallel: the BN gradients interleave with the conv weights in the
            for name in done:                #  arena, so the gradient buckets are released once everything is written)
                after_layer(name)


def fcdensenet_flops_per_image(nodes, ch, h, w):
    """(forward, training) valid-tap FLOPs per image on the LOGICAL channel counts."""
    hw = {"input": (h, w)}
    fwd = train = 0.0
    for n in nodes:
        ih, iw = hw[n.inputs[0]]
        if n.kind == "avgpool":
            hw[n.name] = (ih // 2, iw // 2)
        elif n.kind == "deconv":
            f = deconv_flops(1, ih, iw, ch[n.inputs[0]], n.cout, n.k, n.stride)
            fwd += f; train += 3 * f
            hw[n.name] = (ih * 2, iw * 2)
        else:
            hw[n.name] = (ih, iw)
            if n.kind == "conv":
                f = conv_flops(1, ih, iw, ch[n.inputs[0]], n.cout, n.k, n.k)
                fwd += f; train += (2 if n.inputs[0] == "input" else 3) * f
    return fwd, train
