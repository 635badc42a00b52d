"""Host-side planning for the FCN-8s hot path: layer table, variable arena layout, gradient
buckets and optimizer scalars.  Pure Python (no CUDA) so it is unit-tested on CPU.

Layer/variable order is the creation order of `Network/model/FCN.py:52-107`; names and
layouts are the reference's (`<scope>/weights` HWIO, `<scope>/biases`, `conv_t3/bias`)."""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class Layer:
    name: str
    kind: str          # "conv" | "pool" | "deconv"
    k: int = 0
    cin: int = 0
    cout: int = 0
    stride: int = 1
    relu: bool = True
    bias_name: str = "biases"
    dropout: bool = False   # tf.nn.dropout applied to the output (FCN.py:79,83)

    @property
    def weight_shape(self) -> Tuple[int, int, int, int]:
        if self.kind == "conv":
            return (self.k, self.k, self.cin, self.cout)            # HWIO, FCN.py:125
        return (self.k, self.k, self.cout, self.cin)                # deconv_layer, FCN.py:143

    @property
    def tensor_core(self) -> bool:
        """Layers whose channel counts feed a tcgen05 tile directly (both multiples of 64)."""
        return self.path == "tc"

    @property
    def path(self) -> str:
        """Kernel route: 'tc' implicit GEMM; 'first' = 3x3 conv on the raw image (Cin 1/3/4) with the
        patches built in shared memory; 'im2col' = other layers with kh*kw*Cin <= 64 turned into
        a 1x1 GEMM over a 64-wide patch tensor; 'packed' = transposed conv with tiny Cout as ONE 4-tap
        phase-packed GEMM over the block grid (conv_t3); 'patch' = the same layer as GEMMs in patch space
        [N,H,W,k*k*Cout] (shapes the packed form does not take); 'small' = CUDA-core kernels."""
        if self.kind == "pool":
            return "none"
        if self.kind == "conv":
            if self.cin % 64 == 0 and self.cout % 64 == 0:
                return "tc"
            if self.k == 3 and self.cin in (1, 3, 4) and self.cout in (64, 128, 256):
                return "first"
            if self.k * self.k * self.cin <= 64 and self.cout % 64 == 0:
                return "im2col"
            return "small"
        if self.cin % 64 == 0 and self.cout % 64 == 0 and self.k == 4 and self.stride == 2:
            return "tc"
        if (self.cin % 64 == 0 and self.k == 2 * self.stride and self.stride % 2 == 0 and self.cout % 2 == 0
                and self.stride * self.stride * self.cout in (64, 128, 256) and 32 % (self.stride * self.cout) == 0):
            return "packed"
        if self.cin % 64 == 0 and (self.k * self.k * self.cout) % 64 == 0 and self.k == 2 * self.stride:
            return "patch"
        return "small"


def fcn8s_layers(cin: int = 3, ncls: int = 2, fc: int = 4096) -> List[Layer]:
    """The FCN.py:52-107 graph as a flat list (pool5 -> conv6 -> ... -> conv_t3)."""
    L = []
    def conv(name, k, ci, co, **kw):
        L.append(Layer(name, "conv", k, ci, co, **kw))
    conv("conv1_1", 3, cin, 64); conv("conv1_2", 3, 64, 64); L.append(Layer("pool1", "pool", cout=64))
    conv("conv2_1", 3, 64, 128); conv("conv2_2", 3, 128, 128); L.append(Layer("pool2", "pool", cout=128))
    conv("conv3_1", 3, 128, 256); conv("conv3_2", 3, 256, 256); conv("conv3_3", 3, 256, 256)
    L.append(Layer("pool3", "pool", cout=256))
    conv("conv4_1", 3, 256, 512); conv("conv4_2", 3, 512, 512); conv("conv4_3", 3, 512, 512)
    conv("conv4_4", 3, 512, 512); L.append(Layer("pool4", "pool", cout=512))
    conv("conv5_1", 3, 512, 512); conv("conv5_2", 3, 512, 512); conv("conv5_3", 3, 512, 512)
    L.append(Layer("pool5", "pool", cout=512))
    conv("conv6", 7, 512, fc, dropout=True)
    conv("conv7", 1, fc, fc, dropout=True)
    conv("conv8", 1, fc, ncls)                                       # ReLU'd, FCN.py:86 via :134
    L.append(Layer("conv_t1", "deconv", 4, ncls, 512, stride=2, relu=False))
    L.append(Layer("conv_t2", "deconv", 4, 512, 256, stride=2, relu=False))
    L.append(Layer("conv_t3", "deconv", 16, 256, ncls, stride=8, relu=False, bias_name="bias"))
    return L


def variable_shapes(cin: int = 3, ncls: int = 2, fc: int = 4096) -> "OrderedDict[str, tuple]":
    shapes = OrderedDict()
    for l in fcn8s_layers(cin, ncls, fc):
        if l.kind == "pool":
            continue
        shapes[f"{l.name}/weights"] = l.weight_shape
        shapes[f"{l.name}/{l.bias_name}"] = (l.cout,)
    return shapes


@dataclass(frozen=True)
class Slot:
    name: str
    shape: tuple
    offset: int   # in fp32 elements, multiple of ALIGN
    size: int


ALIGN = 64  # elements: every variable starts 256-byte aligned inside the arena


def arena_layout(shapes) -> Tuple["OrderedDict[str, Slot]", int]:
    """Pack variables (creation order) into one flat fp32 arena; returns (slots, total)."""
    slots = OrderedDict()
    off = 0
    for name, shape in shapes.items():
        size = int(math.prod(shape))
        slots[name] = Slot(name, tuple(shape), off, size)
        off += -(-size // ALIGN) * ALIGN
    return slots, off


def gradient_buckets(slots, layer_groups=None) -> List[Tuple[int, int, str]]:
    """Contiguous arena ranges (start, end, last_layer) in BACKWARD completion order.

    Backward visits layers in reverse creation order, so a bucket is a suffix-to-prefix
    slice of the arena and is complete once `last_layer` (its first layer in creation
    order) has produced its weight gradient.  Default groups follow SURVEY §8e:
    {conv_t3..conv7}, {conv6}, {conv5_x, conv4_x}, {conv3_x..conv1_x}."""
    if layer_groups is None:
        layer_groups = [("conv7", None), ("conv6", "conv7"), ("conv4_1", "conv6"), (None, "conv4_1")]
    names = list(slots)
    def first_index(layer):
        for i, n in enumerate(names):
            if n.split("/")[0] == layer:
                return i
        raise KeyError(layer)
    total_end = max(s.offset + -(-s.size // ALIGN) * ALIGN for s in slots.values())
    out = []
    for start_layer, end_layer in layer_groups:
        i0 = 0 if start_layer is None else first_index(start_layer)
        start = slots[names[i0]].offset
        end = total_end if end_layer is None else slots[names[first_index(end_layer)]].offset
        if end > start:
            out.append((start, end, names[i0].split("/")[0]))
    return out


def gradient_buckets_even(slots, layer_names, nbuckets: int = 4) -> List[Tuple[int, int, str]]:
    """Generic variant for any layer list (creation order): ~equal-size contiguous arena slices cut at
    layer boundaries, returned in backward completion order with the layer that completes each."""
    starts = []
    for ln in layer_names:
        key = f"{ln}/weights"
        if key in slots:
            starts.append((slots[key].offset, ln))
    total_end = max(s.offset + -(-s.size // ALIGN) * ALIGN for s in slots.values())
    target = total_end / float(max(nbuckets, 1))
    out = []
    end = total_end
    for off, ln in reversed(starts):
        if end - off >= target or off == starts[0][0]:
            out.append((off, end, ln))
            end = off
    if out and out[-1][0] != 0:            # variables created before the first listed layer
        out[-1] = (0, out[-1][1], out[-1][2])
    return out


def adam_lr_t(lr: float, t: int, beta1: float = 0.9, beta2: float = 0.999) -> float:
    """tf.train.AdamOptimizer's per-step rate: lr*sqrt(1-b2^t)/(1-b1^t), t starts at 1."""
    return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Batch-sharded data parallelism: rank r owns images [lo, hi) of the global batch."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def train_flops_per_image(h: int, w: int, cin: int = 3, ncls: int = 2, fc: int = 4096, valid_taps: bool = True):
    """(forward, training) FLOPs per image.  Valid-tap count excludes multiplies against SAME
    zero padding; training = fwd + dgrad + wgrad, no dgrad for conv1_1."""
    def valid_frac(n, k):   # average fraction of taps inside [0,n) over output positions, 1-D
        p = k // 2
        tot = 0
        for o in range(n):
            lo, hi = max(o - p, 0), min(o + k - 1 - p, n - 1)
            tot += hi - lo + 1
        return tot / (n * k)
    fwd = 0.0
    train = 0.0
    ch, cw = h, w
    for l in fcn8s_layers(cin, ncls, fc):
        if l.kind == "pool":
            ch, cw = ch // 2, cw // 2
            continue
        if l.kind == "conv":
            f = 2.0 * ch * cw * l.k * l.k * l.cin * l.cout
            if valid_taps:
                f *= valid_frac(ch, l.k) * valid_frac(cw, l.k)
        else:
            # scatter form: every input pixel times k*k taps; valid fraction = (s*in)/( (in-1)*s + k ) per dim
            f = 2.0 * ch * cw * l.k * l.k * l.cin * l.cout
            if valid_taps:
                s = l.stride
                f *= (s * ch) / ((ch - 1) * s + l.k) * (s * cw) / ((cw - 1) * s + l.k)
            ch, cw = ch * l.stride, cw * l.stride
        fwd += f
        train += f * (2.0 if l.name == "conv1_1" else 3.0)
    return fwd, train
