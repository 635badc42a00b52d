"""Encoder-decoder builders on the shared layer helpers of the reference (`Network/utils/utils.py`),
executed on the same B200 kernels as FCN-8s.

Helpers mirrored (same names / argument meaning, graph-building style):
    Conv2D_Block(x, num_filters, 3, 3, batch_normalization=, relu=, name=)     utils.py:185-208 (no bias, :180)
    Deconv2D_Block(x, shape, num_filters, output_shape, 4, 4, 2, name=)         utils.py:255-298 (no bias)
    Batch_Normalization  -> inference-mode affine, gamma/beta trainable           utils.py:300-301
    Max_Pooling(x, name) / Concat([a, b], axis=-1, name)                         utils.py:306,332

Builders:
    UNet(x, num_classes)   BASELINE configs[2].  The reference has no U-Net file; per SURVEY §8a row 12
                           it is SegNet's VGG16-BN encoder (SegNet.py:30-51) with the skip-concat decoder
                           idiom of FCDenseNet (FCDenseNet.py:141-160): Deconv2D_Block(4x4, s2) ->
                           Concat(skip) -> Conv2D_Block..., ReLU on, 1x1 `final_conv` head.

A builder returns a GraphNet with the FCN-compatible training interface (forward / loss / backward /
vars / feed), so AdamOptimizer.minimize(net) and the data-parallel all-reduce work unchanged."""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import plan as P
from .fcn import _Feed, _Feeds
from .ops import Ops, conv_flops
from .overlap import SideStream

BN_EPS = 1e-3          # tf.layers.batch_normalization default epsilon
BN_SCALE = 1.0 / math.sqrt(1.0 + BN_EPS)   # moving_variance stays 1, moving_mean 0 (never updated)


@dataclass
class Node:
    name: str
    kind: str                       # "conv" | "deconv" | "pool" | "concat"
    inputs: List[str] = field(default_factory=list)
    k: int = 3
    cout: int = 0
    stride: int = 1
    bias: bool = False
    bn: bool = False
    relu: bool = False
    bn_scope: str = ""              # TF variable scope of the BN layer ("batch_normalization_7")


class GraphBuilder:
    """Collects helper calls (utils.py signatures) into a node list."""

    def __init__(self):
        self.nodes: List[Node] = []
        self._bn = 0

    def _bn_scope(self):
        s = "batch_normalization" if self._bn == 0 else f"batch_normalization_{self._bn}"
        self._bn += 1
        return s

    def Conv2D_Block(self, x, num_filters, filter_height=3, filter_width=3, batch_normalization=False, relu=False,
                     name=None):
        assert filter_height == filter_width
        n = Node(name, "conv", [x], k=filter_height, cout=num_filters, bn=batch_normalization, relu=relu)
        if batch_normalization:
            n.bn_scope = self._bn_scope()
        self.nodes.append(n)
        return name

    def Deconv2D_Block(self, x, num_filters_out, filter_height=4, filter_width=4, stride=2, name=None):
        self.nodes.append(Node(name, "deconv", [x], k=filter_height, cout=num_filters_out, stride=stride))
        return name

    def Max_Pooling(self, x, name):
        self.nodes.append(Node(name, "pool", [x]))
        return name

    def Concat(self, xs, name):
        self.nodes.append(Node(name, "concat", list(xs)))
        return name


def unet_nodes(num_classes=2):
    """SegNet.py:30-51 encoder (ReLU on) + FCDenseNet.py:141-160 decoder idiom; see module docstring."""
    g = GraphBuilder()
    c = lambda x, f, name: g.Conv2D_Block(x, f, batch_normalization=True, relu=True, name=name)
    x = "input"
    conv1 = c(x, 64, "conv1"); conv2 = c(conv1, 64, "conv2"); pool1 = g.Max_Pooling(conv2, "pool1")
    conv3 = c(pool1, 128, "conv3"); conv4 = c(conv3, 128, "conv4"); pool2 = g.Max_Pooling(conv4, "pool2")
    conv5 = c(pool2, 256, "conv5"); conv6 = c(conv5, 256, "conv6"); conv7 = c(conv6, 256, "conv7")
    pool3 = g.Max_Pooling(conv7, "pool3")
    conv8 = c(pool3, 512, "conv8"); conv9 = c(conv8, 512, "conv9"); conv10 = c(conv9, 512, "conv10")
    pool4 = g.Max_Pooling(conv10, "pool4")
    conv11 = c(pool4, 512, "conv11"); conv12 = c(conv11, 512, "conv12"); conv13 = c(conv12, 512, "conv13")
    pool5 = g.Max_Pooling(conv13, "pool5")
    up1 = g.Deconv2D_Block(pool5, 512, name="unpool1"); cat1 = g.Concat([up1, conv13], "concat1")
    conv14 = c(cat1, 512, "conv14"); conv15 = c(conv14, 512, "conv15"); conv16 = c(conv15, 512, "conv16")
    up2 = g.Deconv2D_Block(conv16, 512, name="unpool2"); cat2 = g.Concat([up2, conv10], "concat2")
    conv17 = c(cat2, 512, "conv17"); conv18 = c(conv17, 512, "conv18"); conv19 = c(conv18, 256, "conv19")
    up3 = g.Deconv2D_Block(conv19, 256, name="unpool3"); cat3 = g.Concat([up3, conv7], "concat3")
    conv20 = c(cat3, 256, "conv20"); conv21 = c(conv20, 256, "conv21"); conv22 = c(conv21, 128, "conv22")
    up4 = g.Deconv2D_Block(conv22, 128, name="unpool4"); cat4 = g.Concat([up4, conv4], "concat4")
    conv23 = c(cat4, 128, "conv23"); conv24 = c(conv23, 64, "conv24")
    up5 = g.Deconv2D_Block(conv24, 64, name="unpool5"); cat5 = g.Concat([up5, conv2], "concat5")
    conv25 = c(cat5, 64, "conv25")
    g.Conv2D_Block(conv25, num_classes, 1, 1, name="final_conv")          # FCDenseNet.py:157
    return g.nodes


def segnet_nodes(num_classes=2):
    """The reference's SegNet as written (SegNet.py:28-87): Conv2D_Block(batch_normalization=True) with the
    helper's default relu=False (utils.py:194), 4x4 s2 Deconv2D_Block "unpool" layers without skips, and a
    3x3 Conv2D_Layer to num_classes + Batch_Normalization (SegNet.py:80-81)."""
    g = GraphBuilder()
    c = lambda x, f, name: g.Conv2D_Block(x, f, batch_normalization=True, name=name)
    x = c("input", 64, "conv1"); x = c(x, 64, "conv2"); x = g.Max_Pooling(x, "pool1")
    x = c(x, 128, "conv3"); x = c(x, 128, "conv4"); x = g.Max_Pooling(x, "pool2")
    x = c(x, 256, "conv5"); x = c(x, 256, "conv6"); x = c(x, 256, "conv7"); x = g.Max_Pooling(x, "pool3")
    x = c(x, 512, "conv8"); x = c(x, 512, "conv9"); x = c(x, 512, "conv10"); x = g.Max_Pooling(x, "pool4")
    x = c(x, 512, "conv11"); x = c(x, 512, "conv12"); x = c(x, 512, "conv13"); x = g.Max_Pooling(x, "pool5")
    x = g.Deconv2D_Block(x, 512, name="unpool1"); x = c(x, 512, "conv14"); x = c(x, 512, "conv15"); x = c(x, 512, "conv16")
    x = g.Deconv2D_Block(x, 512, name="unpool2"); x = c(x, 512, "conv17"); x = c(x, 512, "conv18"); x = c(x, 256, "conv19")
    x = g.Deconv2D_Block(x, 256, name="unpool3"); x = c(x, 256, "conv20"); x = c(x, 256, "conv21"); x = c(x, 128, "conv22")
    x = g.Deconv2D_Block(x, 128, name="unpool4"); x = c(x, 128, "conv23"); x = c(x, 64, "conv24")
    x = g.Deconv2D_Block(x, 64, name="unpool5"); x = c(x, 64, "conv25")
    g.Conv2D_Block(x, num_classes, 3, 3, batch_normalization=True, name="conv26")     # Conv2D_Layer + Batch_Normalization
    return g.nodes


def graph_variable_shapes(nodes, cin):
    """Ordered {name: shape}: `<scope>/weights` (HWIO; deconv [k,k,Cout,Cin]), BN gamma/beta in TF naming."""
    ch = {"input": cin}
    shapes = OrderedDict()
    for n in nodes:
        if n.kind == "pool":
            ch[n.name] = ch[n.inputs[0]]
        elif n.kind == "concat":
            ch[n.name] = sum(ch[i] for i in n.inputs)
        else:
            ci = ch[n.inputs[0]]
            shapes[f"{n.name}/weights"] = (n.k, n.k, ci, n.cout) if n.kind == "conv" else (n.k, n.k, n.cout, ci)
            if n.bias:
                shapes[f"{n.name}/biases"] = (n.cout,)
            if n.bn:
                shapes[f"{n.bn_scope}/gamma"] = (n.cout,)
                shapes[f"{n.bn_scope}/beta"] = (n.cout,)
            ch[n.name] = n.cout
    return shapes


def graph_init(shapes, seed=1234, init="ref"):
    """weights N(0, 0.01^2) (utils.py:179,266), gamma 1, beta/biases 0; 'he' for visibility tests."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("weights"):
            z = rng.standard_normal(shape, dtype=np.float32)
            if init == "ref":
                std = 0.01
            else:
                fan_in = shape[0] * shape[1] * shape[2] if not name.startswith("unpool") else 4 * shape[3]
                std = float(np.sqrt(2.0 / fan_in))
            out[name] = (z * np.float32(std)).astype(np.float32)
        elif name.endswith("gamma"):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


def concat_slots(nodes, shape, route):
    """Zero-copy Concat (utils.py:332): {tensor: (concat node, channel offset)} for the Concat inputs whose producer can write
    straight into its channel slice of the concat buffer and whose gradient can be read in place from the concat's
    gradient (channel-slice views, segk_set_pitch) -- no copy forward, none backward.  That holds for
      * a transposed conv without ReLU that only the Concat reads (the decoder's upsampled half), and
      * a tensor-core conv read by the Concat and by ONE earlier Max_Pooling (an encoder skip): the pool's backward adds
        its gradient to the slice and applies the conv's ReluGrad to the sum; or, without ReLU, read by the Concat alone.
    Everything else keeps the segk_channel_copy form.  `shape`: {tensor: (N,H,W,C)}, `route`: {conv / deconv node: route}."""
    slots = {}
    by_name = {n.name: n for n in nodes}
    order = {n.name: i for i, n in enumerate(nodes)}
    last = nodes[-1].name
    for n in nodes:
        if n.kind != "concat" or n.name == last:
            continue
        off = 0
        for t in n.inputs:
            c = shape[t][3]
            p = by_name.get(t)
            users = [m for m in nodes if t in m.inputs]
            ok = (p is not None and p.kind in ("conv", "deconv") and route.get(t) == "tc" and c % 64 == 0 and off % 64 == 0
                  and t not in slots and n.inputs.count(t) == 1 and t != last)
            if ok and p.kind == "deconv":
                ok = not p.relu and len(users) == 1
            elif ok:
                pools = [m for m in users if m.kind == "pool"]
                if len(users) == 2 and len(pools) == 1:
                    ok = order[pools[0].name] < order[n.name]
                else:
                    ok = len(users) == 1 and not p.relu
            if ok:
                slots[t] = (n.name, off)
            off += c
    return slots


class _Vars:
    """Flat fp32 arenas in creation order (params, grads, optimizer slots)."""

    def __init__(self, shapes, device, values):
        self.shapes = shapes
        self.slots, self.total = P.arena_layout(shapes)
        self.p = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.g = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.m = None
        self.v = None
        for name, arr in values.items():
            self.view(self.p, name).copy_(torch.as_tensor(np.asarray(arr), dtype=torch.float32))
        self.wk, self.wd, self.weff = {}, {}, {}
        self._repack = None

    def view(self, arena, name):
        s = self.slots[name]
        return arena[s.offset:s.offset + s.size].view(s.shape)

    def param(self, name):
        return self.view(self.p, name)

    def grad(self, name):
        return self.view(self.g, name)

    def export(self):
        return OrderedDict((n, self.view(self.p, n).detach().cpu().numpy().copy()) for n in self.slots)

    def repack(self, ops, only=None):
        self._repack(ops, only)


class GraphNet(_Feeds):
    def __init__(self, x, num_classes, nodes, variables=None, init="ref", seed=1234, world_size=1, overlap=True,
                 head_on_tensor_cores=True, zero_copy_concat=True):
        self.head_on_tensor_cores = head_on_tensor_cores      # a k x k conv to num_classes as a 64-column tcgen05 tile
        self.zero_copy_concat = zero_copy_concat              # Concat inputs as channel-slice views of the concat buffer
        if not torch.cuda.is_available():
            raise RuntimeError("GraphNet needs a CUDA (sm_100a) device: the segmentation ops have no CPU fallback")
        x = torch.as_tensor(x)
        self.device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.ops = Ops(self.device)
        self.x = self._as_image(x).to(self.device)
        self.num_classes = int(num_classes)
        if not 2 <= self.num_classes <= 64:
            raise ValueError(f"num_classes must be in [2, 64] (got {self.num_classes})")
        self.N, self.H, self.W, self.Cin = self.x.shape
        self.nodes = nodes
        self.by_name = {n.name: n for n in nodes}
        self.world_size = world_size
        self.keep_prob = 1.0
        self.step_count = 0
        self._init_feeds()
        shapes = graph_variable_shapes(nodes, self.Cin)
        values = variables if variables is not None else graph_init(shapes, seed, init)
        self.vars = _Vars(shapes, self.device, values)
        self.vars._repack = self._repack
        self.vars.fold_mult = BN_SCALE
        self._plan()
        self._repack(self.ops)
        # tensor-core conv layers whose Adam update and bf16 repack -- with the BN scale folded in -- run as one kernel
        # (AdamOptimizer.apply_and_repack): {layer: gamma variable or None}
        self.vars.fused_adam_layers = {n.name: (f"{n.bn_scope}/gamma" if n.bn else None) for n in self.nodes
                                       if n.kind == "conv" and self.route[n.name] == "tc"}
        self._ran_forward = False
        self.side = SideStream(self.device, enabled=overlap)
        self.wside = SideStream(self.device, enabled=overlap)    # weight-gradient GEMMs beside the dgrad chain

    # ---- planning ---------------------------------------------------------------------------
    def _route(self, n, cin):
        if n.kind == "deconv":
            if not (cin % 64 == 0 and n.cout % 64 == 0 and n.k == 4 and n.stride == 2):
                raise ValueError(f"{n.name}: transposed conv needs 4x4 s2 with channels % 64 == 0 (no fallback)")
            return "tc"
        if cin % 64 == 0 and n.cout % 64 == 0:
            return "tc"
        if n.k == 3 and cin in (1, 3, 4) and n.cout in (64, 128, 256) and n.inputs[0] == "input":
            return "first"
        if n.k * n.k * cin <= 64 and n.cout % 64 == 0 and n.inputs[0] == "input":
            return "im2col"
        if n.k == 1 and n.cout in (2, 4, 8) and cin % 8 == 0:
            return "small"
        if (n.k == 3 and n.cout in (2, 4, 8) and cin % 64 == 0 and n.name == self.nodes[-1].name
                and self.H * self.W >= 4096 and getattr(self, "head_on_tensor_cores", True)):
            # k x k head to num_classes (SegNet.py:80) on the tensor cores: weights zero-padded to 64 output channels,
            # fp32 logits written by the narrow epilogue (segk_conv2d_fwd_narrow); 2.5 of SegNet's 15 ms on CUDA cores
            return "head_tc"
        if (n.k in (3, 5) and n.cout in (2, 4) and cin % 8 == 0 and 256 % (cin // 8) == 0
                and n.k * n.k * (cin // 8) * (8 * n.cout + 4) * 4 <= 48 * 1024):
            return "small"        # k x k head to num_classes (SegNet.py:80)
        raise ValueError(f"{n.name}: unsupported conv {n.k}x{n.k} {cin}->{n.cout} (no fallback)")

    def _plan(self):
        dev, N, bf = self.device, self.N, torch.bfloat16
        shape = {"input": (N, self.H, self.W, self.Cin)}
        self.route, self.act, self.idx, self.gbuf, self.patch, self.tmp = {}, OrderedDict(), {}, {}, {}, {}
        last = self.nodes[-1].name
        for n in self.nodes:
            n_, h, w, c = shape[n.inputs[0]]
            if n.kind == "pool":
                shape[n.name] = (N, h // 2, w // 2, c)
            elif n.kind == "concat":
                for i in n.inputs:
                    assert shape[i][:3] == (N, h, w), f"{n.name}: concat inputs differ in size"
                shape[n.name] = (N, h, w, sum(shape[i][3] for i in n.inputs))
            elif n.kind == "deconv":
                self.route[n.name] = self._route(n, c)
                shape[n.name] = (N, h * n.stride, w * n.stride, n.cout)
            else:
                self.route[n.name] = self._route(n, c)
                shape[n.name] = (N, h, w, n.cout)
        self.shape = shape
        self.slot = self._concat_slots()
        for n in self.nodes:
            n_, h, w, c = shape[n.inputs[0]]
            if n.kind == "pool":
                self.idx[n.name] = torch.empty(shape[n.name], dtype=torch.uint8, device=dev)
            elif n.kind == "conv" and self.route[n.name] == "im2col":
                self.patch[n.name] = torch.empty((N, h, w, 64), dtype=bf, device=dev)
                self.tmp[n.name] = torch.empty((1, 1, 64, n.cout), dtype=torch.float32, device=dev)
            if n.name in self.slot:
                continue                                   # a channel slice of its Concat's buffer (below)
            dt = torch.float32 if n.name == last else bf
            self.act[n.name] = torch.empty(shape[n.name], dtype=dt, device=dev)
            self.gbuf[n.name] = torch.empty(shape[n.name], dtype=dt, device=dev)
        for t, (cat, off) in self.slot.items():
            c = shape[t][3]
            self.act[t] = self.act[cat][..., off:off + c]
            self.gbuf[t] = self.gbuf[cat][..., off:off + c]
        # 1-bit ReLU masks of the conv outputs that a tensor-core conv reads (its dgrad epilogue applies the producer's
        # ReluGrad): written by the forward epilogue, 1/16 of the bytes of the bf16 activation
        self.bits = {}
        self.use_mask_bits = True
        for n in self.nodes:
            if n.kind == "conv" and n.relu and self.route[n.name] in ("tc", "first") and n.cout % 32 == 0 and n.name != last:
                readers = [m for m in self.nodes if n.name in m.inputs]
                if any(m.kind == "conv" and self.route[m.name] in ("tc", "head_tc") for m in readers):
                    self.bits[n.name] = torch.empty(shape[n.name][:3] + (n.cout // 32,), dtype=torch.int32, device=dev)
        self.logits = self.act[last]
        assert self.logits.shape[3] == self.num_classes
        self.dlogits = self.gbuf[last]
        self.dlogits_bf16 = torch.empty(shape[last], dtype=bf, device=dev)
        npix = N * self.H * self.W
        self.pred_u8 = torch.empty((N, self.H, self.W), dtype=torch.uint8, device=dev)
        self.loss_sum = torch.zeros(2, dtype=torch.float32, device=dev)   # [sum, mean]
        self.xent_ws = self.ops.xent_workspace(npix, dev)
        self.bn_ws = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
        # per-block partial rows of segk_bn_unfold_grads: one for the wgrad stream (tensor-core layers), one for the main
        # stream (first layer)
        max_c = max([n.cout for n in self.nodes if n.kind == "conv"] + [64])
        self.unfold_ws = [self.ops.bn_unfold_workspace(max_c, dev) for _ in range(2)]
        self.labels = torch.zeros((N, self.H, self.W), dtype=torch.uint8, device=dev)
        # class-count head on the tensor cores: weights / bias padded to 64 output channels, the logit gradient padded to a
        # 64-channel bf16 operand, the padded weight gradient
        self.head_pad, self.head_dz, self.head_gw = {}, {}, {}
        for n in self.nodes:
            if n.kind == "conv" and self.route[n.name] == "head_tc":
                cin = shape[n.inputs[0]][3]
                self.head_pad[n.name] = (torch.zeros((n.k, n.k, cin, 64), dtype=torch.float32, device=dev),
                                         torch.zeros(64, dtype=torch.float32, device=dev))
                self.head_dz[n.name] = torch.empty(shape[n.name][:3] + (64,), dtype=bf, device=dev)
                self.head_gw[n.name] = torch.empty((n.k, n.k, cin, 64), dtype=torch.float32, device=dev)

    def _concat_slots(self):
        if not getattr(self, "zero_copy_concat", True):
            return {}
        return concat_slots(self.nodes, self.shape, self.route)

    def _repack(self, ops, only=None):
        V = self.vars
        for n in self.nodes:
            if n.kind not in ("conv", "deconv"):
                continue
            if only is not None and n.name not in only and n.bn_scope not in only:
                continue
            w = V.param(f"{n.name}/weights")
            if n.bn:   # fold gamma / sqrt(1 + eps) into the weights (columns = Cout for HWIO)
                V.weff[n.name] = ops.scale_columns(w, V.param(f"{n.bn_scope}/gamma"), BN_SCALE, V.weff.get(n.name))
                w = V.weff[n.name]
            r = self.route[n.name]
            if n.kind == "deconv":
                V.wk[n.name], V.wd[n.name] = ops.pack_deconv_weights(w, n.stride, V.wk.get(n.name), V.wd.get(n.name))
            elif r == "tc":
                V.wk[n.name], V.wd[n.name] = ops.pack_conv_weights(w, V.wk.get(n.name), V.wd.get(n.name))
            elif r in ("first", "im2col"):
                V.wk[n.name] = ops.pack_im2col_weights(w, V.wk.get(n.name))
            elif r == "head_tc":
                wp, bp = self.head_pad[n.name]
                ops.remap_weights(w, wp, amap=None, bmap=None, to_phys=True)            # [k,k,Cin,Cout] -> [k,k,Cin,64], zero columns
                b = self._bias(n)
                if b is not None:
                    ops.remap_weights(b.view(1, 1, -1), bp.view(1, 1, -1), amap=None, bmap=None, to_phys=True)
                V.wk[n.name], V.wd[n.name] = ops.pack_conv_weights(wp, V.wk.get(n.name), V.wd.get(n.name))
                V.weff[n.name] = w
            else:
                V.weff[n.name] = w

    def _bias(self, n):
        if n.bn:
            return self.vars.param(f"{n.bn_scope}/beta")
        if n.bias:
            return self.vars.param(f"{n.name}/biases")
        return None

    # ---- forward ----------------------------------------------------------------------------------
    def _in(self, name):
        return self.x if name == "input" else self.act[name]

    def _fusable_pool(self, n):
        """The Max_Pooling node whose 2x2 pool can run in the epilogue of tensor-core conv `n`, and whether the pre-pool
        tensor may stay unwritten (no other consumer; backward never reads it: d(gamma) comes from the weight gradient)."""
        if not getattr(self, "fuse_pool", True) or n.kind != "conv" or self.route[n.name] != "tc":
            return None, False
        if n.name in self.bits:        # (a conv whose mask bits another conv's dgrad reads keeps the plain forward call)
            return None, False
        users = [m for m in self.nodes if n.name in m.inputs]
        pools = [m for m in users if m.kind == "pool"]
        if len(pools) != 1 or self.act[n.name].dtype != torch.bfloat16:
            return None, False
        only = len(users) == 1 and not getattr(self, "keep_prepool", False)
        return pools[0], only

    def _bits_of(self, name):
        return self.bits.get(name) if self.use_mask_bits else None

    def forward(self):
        ops, V = self.ops, self.vars
        pooled = set()
        for n in self.nodes:
            out = self.act[n.name]
            if n.kind == "pool":
                if n.name not in pooled:
                    ops.maxpool_fwd(self._in(n.inputs[0]), out, self.idx[n.name])
            elif n.kind == "concat":
                off = 0
                for i in n.inputs:
                    c = self.shape[i][3]
                    if i not in self.slot:            # (slot tensors were written in place by their producers)
                        ops.channel_copy(self.act[i], 0, out, off, c)
                    off += c
            elif n.kind == "deconv":
                ops.deconv2d_fwd(self._in(n.inputs[0]), V.wk[n.name], self._bias(n), out, n.k, n.stride, relu=n.relu)
            else:
                x, r = self._in(n.inputs[0]), self.route[n.name]
                pool, only = self._fusable_pool(n)
                if pool is not None:          # Max_Pooling in the conv's epilogue (segk_conv2d_fwd_pool)
                    ops.conv2d_fwd_pool(x, V.wk[n.name], self._bias(n), out, self.act[pool.name], self.idx[pool.name], n.k, n.k,
                                        relu=n.relu, pool_only=only)
                    pooled.add(pool.name)
                elif r == "tc":
                    ops.conv2d_fwd(x, V.wk[n.name], self._bias(n), out, n.k, n.k, relu=n.relu, relu_bits=self._bits_of(n.name))
                elif r == "first":
                    ops.conv2d_first_fwd(x, V.wk[n.name], self._bias(n), out, n.k, n.k, relu=n.relu,
                                         relu_bits=self._bits_of(n.name))
                elif r == "im2col":
                    P1 = ops.im2col_k64(x, self.patch[n.name], n.k, n.k)
                    ops.conv2d_fwd(P1, V.wk[n.name], self._bias(n), out, 1, 1, relu=n.relu,
                                   flops=conv_flops(self.N, out.shape[1], out.shape[2], x.shape[3], n.cout, n.k, n.k))
                elif r == "head_tc":
                    ops.conv2d_fwd_narrow(x, V.wk[n.name], self.head_pad[n.name][1] if self._bias(n) is not None else None, out,
                                          n.k, n.k, 64, relu=n.relu)
                else:
                    ops.conv2d_small_fwd(x, V.weff[n.name], self._bias(n), out, relu=n.relu)
        self._ran_forward = True
        return self.logits

    def create(self):
        self.forward()
        self.ops.softmax_infer(self.logits, None, None, self.pred_u8)       # tf.argmax, first index on ties
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

    # ---- backward -----------------------------------------------------------------------------------
    def _relu_mask_of(self, name):
        """Activation whose ReluGrad must be applied to a gradient flowing into tensor `name`."""
        if name == "input":
            return None
        p = self.by_name[name]
        return self.act[name] if (p.kind in ("conv", "deconv") and p.relu) else None

    def backward(self, after_layer=None):
        ops, V = self.ops, self.vars
        has = {n.name: False for n in self.nodes}
        last = self.nodes[-1].name
        has[last] = True
        bias_done = set()        # layers whose d(beta) / BiasAddGrad came out of their consumer's dgrad epilogue

        def fused_bias(t):
            """The input gradient a tensor-core dgrad writes into tensor t -- when this is t's only consumer -- is the
            pre-activation gradient of t's producer: its column sums, summed in the dgrad epilogue, are that layer's
            BiasAddGrad / d(beta) (no separate pass over dz)."""
            p = self.by_name.get(t)
            if p is None or p.kind not in ("conv", "deconv") or has[t] or not (p.bn or p.bias):
                return None
            if sum(1 for m in self.nodes if t in m.inputs) != 1 or self.route.get(p.name) not in ("tc", None):
                return None
            if p.kind == "conv" and self.route[p.name] != "tc":
                return None
            if self.act[t].dtype != torch.bfloat16 or p.cout % 32:
                return None
            c8 = p.cout // 8
            if not (256 % c8 == 0 if c8 <= 256 else c8 % 256 == 0):
                return None
            bias_done.add(p.name)
            return V.grad(f"{p.bn_scope}/beta") if p.bn else V.grad(f"{p.name}/biases")
        for n in reversed(self.nodes):
            if not has[n.name]:
                raise RuntimeError(f"{n.name}: output has no consumer gradient")
            G = self.gbuf[n.name]
            if n.kind == "pool":
                t = n.inputs[0]
                mask = self._relu_mask_of(t)
                if mask is not None and not has[t]:
                    # single gradient path into a ReLU output: the pooled tensor is the same mask at 1/4 the bytes
                    ops.maxpool_bwd(G, self.idx[n.name], self.gbuf[t], pooled=self.act[n.name])
                else:
                    ops.maxpool_bwd(G, self.idx[n.name], self.gbuf[t], act=mask, residual=self.gbuf[t] if has[t] else None)
                has[t] = True
                continue
            if n.kind == "concat":
                off = 0
                for t in n.inputs:
                    c = self.shape[t][3]
                    if t not in self.slot:            # (a slot tensor's gradient IS its slice of G; see _concat_slots)
                        ops.channel_copy(G, off, self.gbuf[t], 0, c, mask=self._relu_mask_of(t), accumulate=has[t])
                    has[t] = True
                    off += c
                continue
            t = n.inputs[0]
            x = self._in(t)
            gw = V.grad(f"{n.name}/weights")
            r = self.route[n.name]
            dz = G
            if r == "small" and G.dtype == torch.float32:
                dz = ops.cast_to_bf16(G, self.dlogits_bf16)
            dz_narrow = dz
            if r == "head_tc":            # logit gradient [N,H,W,classes] -> bf16 [N,H,W,64] operand (zero columns)
                dz = ops.pad_channels(G, self.head_dz[n.name])
            # d(gamma) comes out of the weight gradient (segk_bn_unfold_grads below: dgamma = mult * sum_k W dW', no pass over
            # activations); d(beta) is the BiasAddGrad of dz, off the critical path.  The fp32-logit BN of SegNet's
            # head (SegNet.py:80-81) and non-tensor-core routes keep the activation pass.
            gamma_from_dw = n.bn and n.kind == "conv" and r in ("tc", "first") and self.act[n.name].dtype != torch.float32
            if n.name in bias_done and (gamma_from_dw or not n.bn):
                pass                               # d(beta) / BiasAddGrad: column sums of the consumer's dgrad epilogue
            elif n.bn:
                def bn_grads(dz=dz, n=n, G=G, gamma_from_dw=gamma_from_dw):
                    if self.act[n.name].dtype == torch.float32:
                        ops.bn_grads_f32(G, self.act[n.name], V.param(f"{n.bn_scope}/beta"), V.param(f"{n.bn_scope}/gamma"),
                                         V.grad(f"{n.bn_scope}/gamma"), V.grad(f"{n.bn_scope}/beta"), self.bn_ws)
                    elif gamma_from_dw:
                        ops.bias_grad(dz, V.grad(f"{n.bn_scope}/beta"))
                    else:
                        ops.bn_gamma_grad(dz, self.act[n.name], V.param(f"{n.bn_scope}/beta"),
                                          V.param(f"{n.bn_scope}/gamma"), V.grad(f"{n.bn_scope}/gamma"), self.bn_ws,
                                          dbeta=V.grad(f"{n.bn_scope}/beta"))
                self.side.run(bn_grads)
            elif n.bias and r != "first":      # the fused first-layer wgrad also produces the bias gradient
                self.side.run(lambda dz=dz_narrow, n=n: ops.bias_grad(dz, V.grad(f"{n.name}/biases")))
            # weight gradient (of the folded weights; unfold the BN scale afterwards).  Tensor-core ones go to
            # the wgrad stream, ordered after this point, launched behind the layer's dgrad
            wjob, wmark = None, None

            def unfold(gw=gw, n=n, gamma_from_dw=gamma_from_dw):
                if gamma_from_dw:
                    ops.bn_unfold_grads(gw, V.param(f"{n.name}/weights"), V.param(f"{n.bn_scope}/gamma"), BN_SCALE,
                                        V.grad(f"{n.bn_scope}/gamma"), self.unfold_ws[0 if self.route[n.name] == "tc" else 1])
                elif n.bn:
                    ops.scale_columns(gw, V.param(f"{n.bn_scope}/gamma"), BN_SCALE, gw)

            if n.kind == "deconv":
                wjob = lambda x=x, dz=dz, gw=gw, n=n, unfold=unfold: (ops.deconv2d_wgrad(x, dz, gw, n.k, n.stride), unfold())
            elif r == "tc":
                wjob = lambda x=x, dz=dz, gw=gw, n=n, unfold=unfold: (ops.conv2d_wgrad(x, dz, gw, n.k, n.k), unfold())
            elif r == "first":
                ops.conv2d_first_wgrad(x, dz, gw, n.k, n.k,
                                       dbias=V.grad(f"{n.name}/biases") if (n.bias and not n.bn) else None)
                unfold()
            elif r == "im2col":
                def im2col_wgrad(P=self.patch[n.name], dz=dz, t=self.tmp[n.name], gw=gw, n=n, cin=x.shape[3], unfold=unfold):
                    tmp = ops.conv2d_wgrad(P, dz, t, 1, 1,
                                           flops=conv_flops(self.N, dz.shape[1], dz.shape[2], cin, n.cout, n.k, n.k))
                    gw.view(-1).copy_(tmp.view(-1)[:gw.numel()])
                    unfold()
                wjob = im2col_wgrad      # (every tensor-core wgrad on one stream: they share a partial-sum scratch)
            elif r == "head_tc":
                def head_wgrad(x=x, dz=dz, gw=gw, gwp=self.head_gw[n.name], n=n, unfold=unfold):
                    ops.conv2d_wgrad(x, dz, gwp, n.k, n.k, flops=conv_flops(self.N, x.shape[1], x.shape[2], x.shape[3], n.cout, n.k, n.k))
                    ops.remap_weights(gw, gwp, amap=None, bmap=None, to_phys=False)     # first Cout columns -> the variable's gradient
                    unfold()
                wjob = head_wgrad
            else:
                ops.conv2d_small_wgrad(x, dz, gw)
                unfold()
            if wjob is not None:
                wmark = self.wside.mark()
            # input gradient
            if t != "input":
                mask = self._relu_mask_of(t)
                res = self.gbuf[t] if has[t] else None
                dx = self.gbuf[t]
                if n.kind == "deconv":
                    if res is not None:
                        raise NotImplementedError("deconv input with a second consumer")
                    ops.deconv2d_dgrad(dz, V.wd[n.name], dx, n.k, n.stride, relu_mask=mask, colsum=fused_bias(t))
                elif r == "tc":
                    mbits = self._bits_of(t) if mask is not None else None
                    ops.conv2d_dgrad(dz, V.wd[n.name], dx, n.k, n.k, relu_mask=None if mbits is not None else mask,
                                     relu_mask_bits=mbits, residual=res, colsum=fused_bias(t))
                elif r == "head_tc":
                    mbits = self._bits_of(t) if mask is not None else None
                    ops.conv2d_dgrad(dz, V.wd[n.name], dx, n.k, n.k, relu_mask=None if mbits is not None else mask,
                                     relu_mask_bits=mbits, residual=res, colsum=fused_bias(t),
                                     flops=conv_flops(self.N, dx.shape[1], dx.shape[2], dx.shape[3], n.cout, n.k, n.k))
                elif r == "small":
                    if res is not None:
                        raise NotImplementedError("1x1 head input with a second consumer")
                    ops.conv2d_small_dgrad(dz, V.weff[n.name], dx, relu_mask=mask)
                else:
                    raise NotImplementedError("im2col route is for the first layer only")
                has[t] = True
            if wjob is not None:
                self.wside.run(wjob, after=wmark)
            if after_layer is not None:
                after_layer(n.name)

    def infer(self, image=None):
        if image is not None:
            self.feed({self.image: image})
        self.forward()
        prob = torch.empty_like(self.logits)
        mask = torch.empty((self.N, self.H, self.W), dtype=torch.uint8, device=self.device)
        self.ops.softmax_infer(self.logits, prob, mask)
        self.mark_step_end()
        return prob, mask


def graph_flops_per_image(nodes, h, w, cin):
    """(forward, training) valid-tap FLOPs per image; training = fwd + dgrad + wgrad, no dgrad for the first layer."""
    from .ops import deconv_flops
    shape = {"input": (h, w, cin)}
    fwd = train = 0.0
    for n in nodes:
        ih, iw, ic = shape[n.inputs[0]]
        if n.kind == "pool":
            shape[n.name] = (ih // 2, iw // 2, ic)
        elif n.kind == "concat":
            shape[n.name] = (ih, iw, sum(shape[i][2] for i in n.inputs))
        elif n.kind == "deconv":
            f = deconv_flops(1, ih, iw, ic, n.cout, n.k, n.stride)
            fwd += f; train += 3 * f
            shape[n.name] = (ih * n.stride, iw * n.stride, n.cout)
        else:
            f = conv_flops(1, ih, iw, ic, n.cout, n.k, n.k)
            fwd += f; train += (2 if n.inputs[0] == "input" else 3) * f
            shape[n.name] = (ih, iw, n.cout)
    return fwd, train


def SegNet(x, num_classes=2, **kw):
    """`SegNet(x, num_classes)` of SegNet.py:28 on the same kernels."""
    return GraphNet(x, num_classes, segnet_nodes(num_classes), **kw)


def UNet(x, num_classes=2, **kw):
    """U-Net-style builder (see module docstring); same call shape as SegNet(x, num_classes) (SegNet.py:28)."""
    return GraphNet(x, num_classes, unet_nodes(num_classes), **kw)
