"""FCN-8s on the B200 kernels behind the reference's builder API (`Network/model/FCN.py`).

    net = FCN(image, keep_prob, num_classess)      # FCN.py:31-47  (image: NHWC u8/float CUDA tensor)
    pred, logits = net.create()                    # FCN.py:49-114 (runs the forward pass)
    loss = net.loss(annotation)                    # FCN.py:334
    train_step = AdamOptimizer(1e-4).minimize(net) # FCN.py:338-340
    train_step({net.image: x, net.annotation: y, net.keep_probability: 0.8})   # FCN.py:395-398

The TF-1.x graph/session split collapses: `create()` executes eagerly on the current CUDA
stream.  Every FLOP runs in libsegk.so (tcgen05 implicit-GEMM convs, HBM-bound pool / loss /
optimizer kernels); torch only owns the buffers."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import plan as P
from .ops import Ops, conv_flops, deconv_flops
from .overlap import LocalBuckets, SideStream


def reference_init(shapes, seed: int = 1234, init: str = "ref"):
    """Variables as the reference creates them: weights N(0, 0.01^2) (FCN.py:125,143,102),
    biases 0 (FCN.py:127,145,103); one numpy default_rng(seed) stream consumed in creation
    order.  init='he' gives every layer a visible scale for parity tests."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("weights"):
            z = rng.standard_normal(shape, dtype=np.float32)
            if init == "ref":
                std = 0.01
            else:
                if name.startswith("conv_t"):
                    kh, kw, _, ci = shape
                    s = {4: 2, 16: 8}[kh]
                    fan_in = (kh // s) * (kw // s) * ci
                else:
                    kh, kw, ci, _ = shape
                    fan_in = kh * kw * ci
                std = float(np.sqrt(2.0 / fan_in))
            out[name] = (z * np.float32(std)).astype(np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


class Variables:
    """Flat fp32 arenas (params, Adam m/v, gradients) + bf16 kernel-layout weight shadows."""

    def __init__(self, layers, device, values=None, init="ref", seed=1234):
        self.layers = [l for l in layers if l.kind != "pool"]
        shapes = OrderedDict()
        for l in self.layers:
            shapes[f"{l.name}/weights"] = l.weight_shape
            shapes[f"{l.name}/{l.bias_name}"] = (l.cout,)
        self.shapes = shapes
        self.slots, self.total = P.arena_layout(shapes)
        self.device = device
        self.p = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.g = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.m = None   # optimizer slots are created by the optimizer (as TF does)
        self.v = None
        if values is None:
            if init == "device":
                gen = torch.Generator(device=device)
                gen.manual_seed(seed)
                for name, s in self.slots.items():
                    if name.endswith("weights"):
                        self.view(self.p, name).copy_(
                            torch.randn(s.shape, generator=gen, device=device, dtype=torch.float32) * 0.01)
            else:
                values = reference_init(shapes, seed, init)
        if values is not None:
            self.assign(values)
        self.wk = {}
        self.wd = {}
        # conv layers whose Adam update and bf16 repack run as one kernel (AdamOptimizer.apply_and_repack)
        self.fused_adam_layers = {l.name for l in self.layers if l.kind == "conv" and l.path == "tc"}

    def view(self, arena, name):
        s = self.slots[name]
        return arena[s.offset:s.offset + s.size].view(s.shape)

    def assign(self, values):
        for name, arr in values.items():
            self.view(self.p, name).copy_(torch.as_tensor(np.asarray(arr), dtype=torch.float32))

    def param(self, name):
        return self.view(self.p, name)

    def grad(self, name):
        return self.view(self.g, name)

    def export(self):
        """{reference variable name: numpy array} in reference layouts (checkpoint interchange)."""
        return OrderedDict((n, self.view(self.p, n).detach().cpu().numpy().copy()) for n in self.slots)

    def repack(self, ops: Ops, only=None):
        """fp32 masters -> bf16 kernel layouts for the tensor-core layers (after each update);
        `only`: restrict to these layer names (per-bucket repack)."""
        for l in self.layers:
            if only is not None and l.name not in only:
                continue
            w = self.param(f"{l.name}/weights")
            if l.path == "tc" and l.kind == "conv":
                self.wk[l.name], self.wd[l.name] = ops.pack_conv_weights(w, self.wk.get(l.name), self.wd.get(l.name))
            elif l.path == "tc":
                self.wk[l.name], self.wd[l.name] = ops.pack_deconv_weights(w, l.stride, self.wk.get(l.name),
                                                                           self.wd.get(l.name))
            elif l.path in ("first", "im2col"):
                self.wk[l.name] = ops.pack_im2col_weights(w, self.wk.get(l.name))
            elif l.path == "packed":
                self.wk[l.name], self.wd[l.name] = ops.pack_deconv_packed(w, l.stride, self.wk.get(l.name), self.wd.get(l.name))
            elif l.path == "patch":
                # W[k,k,Cout,Cin] as the matrix [(ky,kx,co)][ci]: wk = fwd B operand, wd = its transpose
                e = l.k * l.k * l.cout
                self.wk[l.name], self.wd[l.name] = ops.pack_matrix(w.view(1, e, l.cin), self.wk.get(l.name),
                                                                   self.wd.get(l.name))


class _Feed:
    """Placeholder handle (tf.placeholder analogue, FCN.py:311-313) used as feed_dict key."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"<placeholder {self.name}>"


class _Feeds:
    """Placeholder feeds shared by FCN and graph.GraphNet: image / annotation / keep_probability
    (FCN.py:311-313,395).  Needs self.device, self.ops, self.N/H/W/Cin, self.num_classes."""

    def _init_feeds(self):
        self._stage_bufs = {}
        self._staged_now = []
        self._copy_stream = None
        self.image = _Feed("input_image")
        self.annotation = _Feed("annotation")
        self.keep_probability = _Feed("keep_probability")

    def _as_image(self, x, device=None):
        """The image feed (FCN.py:312,395).  u8 pixels go to conv1_1 as they are (its kernel reads u8); a
        float32 / bf16 feed -- legal for the reference's f32 placeholder, e.g. mean-subtracted or [0,1]
        images -- is cast to bf16 by `segk_cast_to_bf16` and takes the bf16 first-layer path (the compute
        dtype of this path; raw 0..255 integers are exact in bf16).  Nothing is rounded or clamped."""
        x = torch.as_tensor(x)
        if x.dim() != 4:
            raise ValueError("image must be NHWC")
        if x.dtype == torch.uint8 or x.dtype == torch.bfloat16:
            return x.contiguous()
        if x.dtype != torch.float32:
            raise TypeError(f"image dtype {x.dtype} not supported: feed uint8, float32 or bfloat16")
        device = device if device is not None else (x.device if x.is_cuda else self.device)
        x = x.contiguous().to(device, non_blocking=True)
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=device)
        with torch.cuda.device(device):
            self.ops.cast_to_bf16(x, out)
        return out

    def feed(self, feed_dict):
        for k, v in feed_dict.items():
            name = k.name if isinstance(k, _Feed) else str(k)
            if name == "input_image":
                x = self._as_image(v)
                if tuple(x.shape) != (self.N, self.H, self.W, self.Cin):
                    raise ValueError(f"image shape {tuple(x.shape)} != planned {(self.N, self.H, self.W, self.Cin)}")
                self.x = self._stage(x, "x")
            elif name == "annotation":
                self.labels = self._as_labels(v)
            elif name == "keep_probability":
                self.keep_prob = float(v)
            else:
                raise KeyError(name)

    def _stage(self, t, slot):
        """Host tensors are copied into preallocated device buffers on a COPY stream (two buffers per feed
        slot): the H2D copy of step i+1 is enqueued while the GPU still runs step i and overlaps it; the
        compute stream only waits for the copy's event.  A buffer is reused two steps later, after the
        step-end event of its last user.  Device tensors are used in place."""
        if t.is_cuda:
            return t
        st = self._stage_bufs.setdefault((slot, t.dtype, tuple(t.shape)), {"buf": [None, None], "busy": [None, None], "i": 0})
        i = st["i"] = st["i"] ^ 1
        if st["buf"][i] is None:
            st["buf"][i] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        if st["busy"][i] is not None:
            cs.wait_event(st["busy"][i])
        with torch.cuda.stream(cs):
            st["buf"][i].copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        torch.cuda.current_stream().wait_event(ev)
        self._staged_now.append((st, i))
        return st["buf"][i]

    def mark_step_end(self):
        """Called once the step that consumed the staged feeds is fully enqueued on the current stream."""
        if self._staged_now:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for st, i in self._staged_now:
                st["busy"][i] = ev
            self._staged_now = []

    def _as_labels(self, y):
        """Class-id map u8 [N,H,W] (or [N,H,W,1], the sparse variant of FCN.py:329); also accepts the
        reference's one-hot [N,H,W,num_classes] bool / u8 / float32 annotation (channel 1 = road,
        FCN.py:195-201,313), converted to class ids by `segk_onehot_to_ids`.  Label values >= num_classes
        are ignored by the loss kernel (no loss, zero gradient)."""
        y = torch.as_tensor(y)
        shape = (self.N, self.H, self.W)
        if y.dim() == 4 and y.shape[3] == 1:
            y = y[..., 0]
        if y.dim() == 4:
            if tuple(y.shape) != shape + (self.num_classes,):
                raise ValueError(f"annotation shape {tuple(y.shape)}: expected {shape}, {shape + (1,)} or one-hot {shape + (self.num_classes,)}")
            if y.dtype not in (torch.bool, torch.uint8, torch.float32):
                raise TypeError(f"one-hot annotation dtype {y.dtype} not supported: feed bool, uint8 or float32")
            y = self._stage(y.contiguous(), "onehot")
            ids = torch.empty(shape, dtype=torch.uint8, device=self.device)
            self.ops.onehot_to_ids(y, ids)
            return ids
        if tuple(y.shape) != shape:
            raise ValueError(f"annotation shape {tuple(y.shape)} != {shape}")
        if y.dtype != torch.uint8:
            if y.dtype not in (torch.int32, torch.int64, torch.int16, torch.int8, torch.bool):
                raise TypeError(f"class-id annotation dtype {y.dtype} not supported: feed an integer type")
            y = y.to(torch.uint8)           # dtype conversion of ids only (values 0..C-1)
        return self._stage(y.contiguous(), "labels")



class FCN(_Feeds):
    """FCN-8s builder with the reference's constructor signature (FCN.py:31-47)."""

    def __init__(self, x, keep_prob=1.0, num_classess=2, variables=None, init="ref", seed=1234, fc=4096,
                 dropout_seed=42, world_size=1, overlap=True):
        if not torch.cuda.is_available():
            raise RuntimeError("FCN needs a CUDA (sm_100a) device: the segmentation ops have no CPU fallback")
        x = torch.as_tensor(x)
        self.device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.ops = Ops(self.device)
        x = self._as_image(x).to(self.device)
        self.num_classes = int(num_classess)
        if not 2 <= self.num_classes <= 64:
            raise ValueError(f"num_classess must be in [2, 64] (got {self.num_classes})")
        self.keep_prob = float(keep_prob)
        self.N, self.H, self.W, self.Cin = x.shape
        if self.H % 32 or self.W % 32:
            raise ValueError(f"image size {self.H}x{self.W} must be a multiple of 32 (five 2x2 pools)")
        self.layers = P.fcn8s_layers(self.Cin, self.num_classes, fc)
        self.vars = variables if isinstance(variables, Variables) else Variables(
            self.layers, self.device, values=variables, init=init, seed=seed)
        self.vars.repack(self.ops)
        self.world_size = world_size
        self._init_feeds()
        self.dropout_seed = dropout_seed
        self.step_count = 0
        self.injected_masks = None     # {"dropout6": u8 tensor, "dropout7": ...} for parity runs
        self.fuse_pool = True          # max_pool in the epilogue of the conv in front of it (segk_conv2d_fwd_pool)
        self.keep_prepool = False      # True: also store the pre-pool conv outputs (per-activation parity tests)
        self.use_mask_bits = True      # ReluGrad masks as 1-bit words written by the forward epilogues (self.bits)
        self.x = x
        self._alloc()
        self._ran_forward = False
        self.side = SideStream(self.device, enabled=overlap)
        self.wside = SideStream(self.device, enabled=overlap)      # weight-gradient GEMMs beside the dgrad chain

    # -- buffers ------------------------------------------------------------------------
    def _alloc(self):
        dev, N = self.device, self.N
        bf = torch.bfloat16
        self.act = OrderedDict()
        self.idx = {}
        h, w = self.H, self.W
        for l in self.layers:
            if l.kind == "pool":
                h, w = h // 2, w // 2
                self.act[l.name] = torch.empty((N, h, w, l.cout), dtype=bf, device=dev)
                self.idx[l.name] = torch.empty((N, h, w, l.cout), dtype=torch.uint8, device=dev)
            elif l.kind == "conv":
                self.act[l.name] = torch.empty((N, h, w, l.cout), dtype=bf, device=dev)
            else:
                h, w = h * l.stride, w * l.stride
                if l.name == "conv_t3":
                    self.act[l.name] = torch.empty((N, h, w, l.cout), dtype=torch.float32, device=dev)
                else:
                    self.act[l.name] = torch.empty((N, h, w, l.cout), dtype=bf, device=dev)
        self.patch = {}     # layer -> bf16 patch tensor saved for backward; fp32 patch-space scratch
        self.patch_f32 = {}
        h, w = self.H, self.W
        for l in self.layers:
            if l.kind == "pool":
                h, w = h // 2, w // 2
            elif l.path == "im2col":
                self.patch[l.name] = torch.empty((N, h, w, 64), dtype=bf, device=dev)
                self.patch_f32[l.name] = torch.empty((1, 1, 64, l.cout), dtype=torch.float32, device=dev)
            elif l.kind == "deconv":
                if l.path == "packed":
                    r = l.stride * l.stride * l.cout
                    # dy re-blocked per 2x2 input neighbourhood; zero-initialised once (every entry is rewritten
                    # each step, the ones outside the image with zeros)
                    self.patch[l.name] = torch.zeros((N, h + 1, w + 1, r), dtype=bf, device=dev)
                    self.patch_f32[l.name] = torch.empty((4, l.cin, r), dtype=torch.float32, device=dev)
                elif l.path == "patch":
                    e = l.k * l.k * l.cout
                    self.patch[l.name] = torch.empty((N, h, w, e), dtype=bf, device=dev)
                    self.patch_f32[l.name] = torch.empty((N, h, w, e), dtype=torch.float32, device=dev)
                h, w = h * l.stride, w * l.stride
        # 1-bit ReLU masks (u32 [N,H,W,C/32]) of the conv outputs whose ReluGrad runs in the epilogue of the NEXT conv's
        # dgrad: written by the forward epilogue, read there instead of the bf16 activation (1/16 of the bytes)
        self.bits = {}
        h, w = self.H, self.W
        for i, l in enumerate(self.layers):
            if l.kind == "pool":
                h, w = h // 2, w // 2
            nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
            if (l.kind == "conv" and l.path in ("tc", "first") and l.relu and not l.dropout and l.cout % 32 == 0 and
                    nxt is not None and nxt.kind == "conv" and nxt.path == "tc"):
                self.bits[l.name] = torch.empty((N, h, w, l.cout // 32), dtype=torch.int32, device=dev)
        self.logits = self.act["conv_t3"]
        npix = N * self.H * self.W
        self.dlogits = torch.empty_like(self.logits)
        self.pred_u8 = torch.empty((N, self.H, self.W), dtype=torch.uint8, device=dev)
        self.loss_sum = torch.zeros(2, dtype=torch.float32, device=dev)   # [sum, mean]
        self.cm = torch.zeros(4, dtype=torch.int64, device=dev)
        self.xent_ws = self.ops.xent_workspace(npix, dev)
        self.labels = torch.zeros((N, self.H, self.W), dtype=torch.uint8, device=dev)
        # rotating gradient buffers (largest activation: N*H*W*64 bf16) + the two skip gradients; three,
        # so that a layer's wgrad on the side stream may still read dz while the next dgrad writes
        big = N * self.H * self.W * 64
        self._gbuf = [torch.empty(big, dtype=bf, device=dev) for _ in range(3)]
        self.dfuse_1 = torch.empty_like(self.act["conv_t1"])
        self.dfuse_2 = torch.empty_like(self.act["conv_t2"])

    def _g(self, i, like):
        return self._gbuf[i][:like.numel()].view(like.shape)

    # -- forward (FCN.py:49-114) ------------------------------------------------------
    def _dropout_fwd(self, l, t):
        if self.keep_prob >= 1.0:
            return
        mask = None if self.injected_masks is None else self.injected_masks.get("dropout" + l.name[-1])
        seed = (self.dropout_seed * 1000003 + self.step_count) * 16 + int(l.name[-1])
        self.ops.dropout(t, t, self.keep_prob, seed, mask)

    def _bits_of(self, l):
        return self.bits.get(l.name) if self.use_mask_bits else None

    def create(self):
        """Run the forward pass; returns (pred [N,H,W,1] int64, logits [N,H,W,C] f32)."""
        self.forward()
        # tf.argmax, first index on ties (FCN.py:111), by the softmax kernel; int64 [N,H,W,1] as expand_dims gives
        self.ops.softmax_infer(self.logits, None, None, self.pred_u8)
        pred = self.pred_u8.to(torch.int64).unsqueeze(3)
        return pred, self.logits

    def forward(self):
        ops, V, act = self.ops, self.vars, self.act
        cur = self.x
        L = self.layers
        pooled_by_conv = None
        for i, l in enumerate(L):
            out = act[l.name]
            if l.kind == "pool":
                if pooled_by_conv != l.name:
                    ops.maxpool_fwd(cur, out, self.idx[l.name])
            elif l.kind == "conv":
                b = V.param(f"{l.name}/{l.bias_name}")
                nxt = L[i + 1] if i + 1 < len(L) else None
                if l.path == "tc" and self.fuse_pool and nxt is not None and nxt.kind == "pool" and l.relu and not l.dropout:
                    # conv -> ReLU -> max_pool (FCN.py:54-76): the pool runs in the conv's epilogue; nothing reads the
                    # pre-pool tensor again (MaxPoolGrad routes by the stored index, ReluGrad masks by the pooled value)
                    ops.conv2d_fwd_pool(cur, V.wk[l.name], b, out, act[nxt.name], self.idx[nxt.name], l.k, l.k, relu=True,
                                        pool_only=not self.keep_prepool)
                    pooled_by_conv = nxt.name
                elif l.path == "tc":
                    ops.conv2d_fwd(cur, V.wk[l.name], b, out, l.k, l.k, relu=l.relu, relu_bits=self._bits_of(l))
                elif l.path == "first":
                    # (measured: the bits in this kernel's epilogue cost 47 us; a streaming segk_relu_bits pass on the side
                    # stream under conv1_2's forward costs more of the step, 8.15 vs 8.07 ms)
                    ops.conv2d_first_fwd(cur, V.wk[l.name], b, out, l.k, l.k, relu=l.relu, relu_bits=self._bits_of(l))
                elif l.path == "im2col":
                    P1 = ops.im2col_k64(cur, self.patch[l.name], l.k, l.k)
                    ops.conv2d_fwd(P1, V.wk[l.name], b, out, 1, 1, relu=l.relu,
                                   flops=conv_flops(self.N, out.shape[1], out.shape[2], l.cin, l.cout, l.k, l.k))
                else:
                    ops.conv2d_small_fwd(cur, V.param(f"{l.name}/weights"), b, out, relu=l.relu)
                if l.dropout:
                    self._dropout_fwd(l, out)
            else:
                b = V.param(f"{l.name}/{l.bias_name}")
                res = {"conv_t1": act["pool4"], "conv_t2": act["pool3"]}.get(l.name)   # fuse, FCN.py:92,96
                if l.path == "tc":
                    ops.deconv2d_fwd(cur, V.wk[l.name], b, out, l.k, l.stride, residual=res)
                elif l.path == "packed":
                    assert res is None and out.dtype == torch.float32
                    ops.deconv2d_packed_fwd(cur, V.wk[l.name], b, out, l.k, l.stride)
                elif l.path == "patch":
                    yp = ops.conv2d_fwd(cur, V.wk[l.name], None, self.patch_f32[l.name], 1, 1, relu=False,
                                        flops=deconv_flops(self.N, cur.shape[1], cur.shape[2], l.cin, l.cout, l.k, l.stride))
                    ops.deconv_col2im(yp, b, out, l.k, l.stride, residual=res)
                else:
                    ops.deconv2d_small_fwd(cur, V.param(f"{l.name}/weights"), b, out, l.stride, residual=res)
            cur = out
        self._ran_forward = True
        return self.logits

    # -- loss (FCN.py:334) ----------------------------------------------------------------
    def loss(self, annotation=None, with_grad=False):
        """reduce_mean(softmax_cross_entropy_with_logits); returns a 0-dim device tensor."""
        if annotation is not None:
            self.labels = self._as_labels(annotation)
        if not self._ran_forward:
            self.forward()
        npix = self.N * self.H * self.W
        self.ops.softmax_xent(self.logits, self.labels, self.dlogits if with_grad else None, self.pred_u8,
                              self.loss_sum, None, self.xent_ws, 1.0 / (npix * self.world_size))
        return self.loss_sum[1]

    def confusion_matrix(self):
        """Road / non-road confusion counts cm[gt, pred] of the last forward (int64 2x2)."""
        cm = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.ops.confusion_matrix(self.labels, self.pred_u8, cm)
        return cm.view(2, 2)

    # -- backward (what optimizer.minimize emits, FCN.py:340; schedule = SURVEY App. E) ---------
    def backward(self, after_layer=None):
        """Fills vars.g with d(mean loss)/d(variable).  Needs forward() + loss(with_grad=True).
        `after_layer(name)` is called once a layer's weight and bias gradients are enqueued
        (the data-parallel path uses it to launch bucket all-reduces)."""
        ops, V, act = self.ops, self.vars, self.act
        L = self.layers
        inv_keep = 1.0 / self.keep_prob if self.keep_prob < 1.0 else 1.0
        names = [l.name for l in L]
        def prev_act(i):
            return self.x if i == 0 else act[names[i - 1]]
        dcur = self.dlogits          # gradient wrt the current layer's output (pre-activation-grad applied)
        bias_done = set()            # layers whose BiasAddGrad came out of another kernel's pass
        flip = 0
        nbuf = len(self._gbuf)

        def next_dx(like):
            nonlocal flip
            dx = self._g(flip, like)
            flip = (flip + 1) % nbuf
            self.side.before_write(dx)
            self.wside.before_write(dx)
            return dx

        def fused_bias(prev):
            """The input gradient a tensor-core dgrad writes is the pre-activation gradient of its producer layer:
            its column sums are that layer's BiasAddGrad, summed in the dgrad epilogue (no second pass over dz)."""
            if prev is None or prev.kind == "pool" or prev.path == "first" or prev.cout % 32:
                return None
            c8 = prev.cout // 8
            if not (256 % c8 == 0 if c8 <= 256 else c8 % 256 == 0):
                return None
            bias_done.add(prev.name)
            return V.grad(f"{prev.name}/{prev.bias_name}")

        for i in range(len(L) - 1, -1, -1):
            l = L[i]
            xin = prev_act(i)
            if l.kind == "pool":
                # MaxPoolGrad fused with the ReluGrad of the pre-pool conv output
                dx = next_dx(xin)
                # the pooled output doubles as the ReluGrad mask of the conv before the pool, and the column sums
                # of the masked gradient are that conv's BiasAddGrad
                before = L[i - 1]
                db = None
                if before.kind == "conv" and before.path != "first" and (before.cout // 8) <= 256 and 256 % (before.cout // 8) == 0:
                    db = V.grad(f"{before.name}/{before.bias_name}")
                    bias_done.add(before.name)
                ops.maxpool_bwd(dcur, self.idx[l.name], dx, pooled=act[l.name], dbias=db)
                dcur = dx
                continue
            gw = V.grad(f"{l.name}/weights")
            gb = V.grad(f"{l.name}/{l.bias_name}")
            # BiasAddGrad only reads dcur: HBM-bound, off the critical path -> side stream
            # (the fused first-layer wgrad produces it as one more row of its GEMM)
            if l.path != "first" and l.name not in bias_done:
                self.side.run(lambda d=dcur, g=gb: ops.bias_grad(d, g), reads=(dcur,))
            prev = L[i - 1] if i > 0 else None
            # tensor-core weight gradients go to the wgrad stream, ordered after this point (dcur is
            # ready) but launched after the layer's dgrad so the critical-path kernel gets the SMs first
            wjob, wmark, wreads = None, None, (dcur,)
            if l.kind == "deconv":
                if l.path == "tc":
                    wjob = lambda x=xin, d=dcur, g=gw, l=l: ops.deconv2d_wgrad(x, d, g, l.k, l.stride)
                    wmark = self.wside.mark()
                elif l.path == "packed":
                    Pg = ops.deconv_pack_dy(dcur, self.patch[l.name], l.stride)
                    wjob = lambda Pg=Pg, x=xin, g=gw, l=l: ops.deconv2d_packed_wgrad(x, Pg, g, self.patch_f32[l.name], l.k, l.stride)
                    wmark = self.wside.mark()
                    wreads = ()
                elif l.path == "patch":
                    Pg = ops.deconv_patch_gather(dcur, self.patch[l.name], l.k, l.stride)
                    dfl = deconv_flops(self.N, xin.shape[1], xin.shape[2], l.cin, l.cout, l.k, l.stride)
                    # (every tensor-core wgrad runs on the wgrad stream: they share one partial-sum scratch)
                    wjob = lambda Pg=Pg, x=xin, g=gw, l=l, dfl=dfl: ops.conv2d_wgrad(
                        Pg, x, g.view(1, 1, l.k * l.k * l.cout, l.cin), 1, 1, flops=dfl)
                    wmark = self.wside.mark()
                    wreads = ()
                else:
                    ops.deconv2d_small_wgrad(xin, dcur, gw, l.stride)
                # gradient wrt the deconv input
                if l.name == "conv_t3":
                    dx = self.dfuse_2
                elif l.name == "conv_t2":
                    dx = self.dfuse_1
                else:
                    dx = next_dx(xin)
                mask = xin if (prev is not None and prev.kind == "conv" and prev.relu) else None
                cs = fused_bias(prev) if l.path in ("tc", "patch", "packed") else None
                if l.path == "tc":
                    ops.deconv2d_dgrad(dcur, V.wd[l.name], dx, l.k, l.stride, relu_mask=mask, colsum=cs)
                elif l.path == "packed":
                    ops.deconv2d_packed_dgrad(Pg, V.wd[l.name], dx, l.cout, l.k, l.stride, relu_mask=mask, colsum=cs)
                elif l.path == "patch":
                    ops.conv2d_dgrad(Pg, V.wd[l.name], dx, 1, 1, relu_mask=mask, flops=dfl, colsum=cs)
                else:
                    ops.deconv2d_small_dgrad(dcur, V.param(f"{l.name}/weights"), dx, l.stride, relu_mask=mask)
                if wjob is not None:
                    self.wside.run(wjob, reads=wreads, after=wmark)
                dcur = dx
            else:
                if l.path == "tc":
                    wjob = lambda x=xin, d=dcur, g=gw, l=l: ops.conv2d_wgrad(x, d, g, l.k, l.k)
                    wmark = self.wside.mark()
                elif l.path == "first":
                    if i != 0:
                        raise NotImplementedError("'first' route is for the first layer only (no input gradient)")
                    ops.conv2d_first_wgrad(xin, dcur, gw, l.k, l.k, dbias=gb)
                elif l.path == "im2col":
                    if i != 0:
                        raise NotImplementedError("im2col route is for the first layer only (no input gradient)")
                    def im2col_wgrad(P=self.patch[l.name], d=dcur, t=self.patch_f32[l.name], g=gw, l=l):
                        tmp = ops.conv2d_wgrad(P, d, t, 1, 1,
                                               flops=conv_flops(self.N, d.shape[1], d.shape[2], l.cin, l.cout, l.k, l.k))
                        g.view(-1).copy_(tmp.view(-1)[:g.numel()])      # rows >= K are the zero padding
                    wjob = im2col_wgrad
                    wmark = self.wside.mark()
                else:
                    ops.conv2d_small_wgrad(xin, dcur, gw)
                dnext = dcur
                if i > 0:
                    # ReluGrad of the producer (mask = its output) and dropout backward (scale) fused
                    mask = xin if (prev.kind == "conv" and prev.relu) else None
                    scale = inv_keep if (prev.kind == "conv" and prev.dropout) else 1.0
                    # AddN of the skip gradients into pool4 / pool3 (FCN.py:92,96)
                    res = {"pool4": self.dfuse_1, "pool3": self.dfuse_2}.get(prev.name)
                    dx = next_dx(xin)
                    if l.path == "tc":
                        mbits = self._bits_of(prev) if mask is not None else None
                        ops.conv2d_dgrad(dcur, V.wd[l.name], dx, l.k, l.k, relu_mask=None if mbits is not None else mask,
                                         relu_mask_bits=mbits, residual=res, scale=scale, colsum=fused_bias(prev))
                    else:
                        assert res is None
                        ops.conv2d_small_dgrad(dcur, V.param(f"{l.name}/weights"), dx, relu_mask=mask, scale=scale)
                    dnext = dx
                if wjob is not None:
                    self.wside.run(wjob, reads=wreads, after=wmark)
                dcur = dnext
            if after_layer is not None:
                after_layer(l.name)

    # -- inference (FCN.py:229,204-206) -----------------------------------------------------
    def infer_graphed(self, image):
        return _infer_graphed(self, image)

    def infer(self, image=None):
        """softmax at keep_prob 1.0 and the road mask softmax[...,1] > 0.5."""
        if image is not None:
            self.feed({self.image: image})
        kp, self.keep_prob = self.keep_prob, 1.0
        mb, self.use_mask_bits = self.use_mask_bits, False        # (no backward follows: no ReluGrad masks)
        try:
            self.forward()
        finally:
            self.keep_prob = kp
            self.use_mask_bits = mb
        prob = torch.empty_like(self.logits)
        mask = torch.empty((self.N, self.H, self.W), dtype=torch.uint8, device=self.device)
        self.ops.softmax_infer(self.logits, prob, mask)
        self.mark_step_end()
        return prob, mask


def _infer_graphed(net, image):
    """`net.infer(image)` replayed from a CUDA graph: the forward pass + softmax + road mask are captured once (the
    launches are static: same pointers, shapes and tensor maps every call) and each call is one host->static-buffer
    copy plus one graph launch -- for the per-image inference loop of gen_test_output (FCN.py:224-231), where the
    23 kernel launches of a batch-1 forward cost more host time than GPU time.  Returns (prob, mask) in static
    buffers that the next call overwrites."""
    st = net.__dict__.setdefault("_igraph", {})
    img = net._as_image(image)
    if tuple(img.shape) != (net.N, net.H, net.W, net.Cin):
        raise ValueError(f"image shape {tuple(img.shape)} != planned {(net.N, net.H, net.W, net.Cin)}")
    if not st:
        st["x"] = torch.empty((net.N, net.H, net.W, net.Cin), dtype=img.dtype, device=net.device)
        st["prob"] = torch.empty_like(net.logits)
        st["mask"] = torch.empty((net.N, net.H, net.W), dtype=torch.uint8, device=net.device)
        st["x"].copy_(img)

        def body():
            kp, net.keep_prob = net.keep_prob, 1.0
            mb = getattr(net, "use_mask_bits", False)
            net.use_mask_bits = False
            x0, net.x = net.x, st["x"]
            try:
                net.forward()
                net.ops.softmax_infer(net.logits, st["prob"], st["mask"])
            finally:
                net.keep_prob, net.use_mask_bits, net.x = kp, mb, x0

        side = torch.cuda.Stream(net.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):            # warm-up outside the capture: the context's grow-only scratch is allocated here
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(net.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        st["graph"] = g
    if st["x"].dtype != img.dtype:
        raise TypeError("graphed inference was captured for another image dtype")
    st["x"].copy_(img, non_blocking=True)
    st["graph"].replay()
    return st["prob"], st["mask"]


def gen_test_output(net, images):
    """gen_test_output of FCN.py:213-233 without the file IO: for each u8 NHWC batch run the softmax at
    keep_prob 1.0, threshold the road probability at 0.5 and paste the [0,255,0,127] overlay
    (FCN.py:229-231).  Yields the overlaid u8 batches (device tensors)."""
    for batch in images:
        prob, mask = net.infer(batch)
        yield net.ops.overlay_mask(net.x, mask)


class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate) with TF's update formula (FCN.py:338)."""

    def __init__(self, learning_rate=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.beta1, self.beta2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0

    def minimize(self, net: FCN, allreduce=None):
        V = net.vars
        if V.m is None:
            V.m = torch.zeros_like(V.p)
            V.v = torch.zeros_like(V.p)
        return TrainStep(net, self, allreduce)

    def apply(self, net: FCN, lo=0, hi=None, grad_scale=1.0):
        V = net.vars
        hi = V.total if hi is None else hi
        lr_t = P.adam_lr_t(self.lr, self.t, self.beta1, self.beta2)
        net.ops.adam_step(V.p[lo:hi], V.m[lo:hi], V.v[lo:hi], V.g[lo:hi], lr_t, self.beta1, self.beta2, self.eps,
                          grad_scale)

    def apply_and_repack(self, net, lo=0, hi=None, names=None, grad_scale=1.0):
        """ApplyAdam on the arena slice [lo, hi) + the bf16 repack of its layers.  The conv weights of the tensor-core
        layers (nearly all parameters) take ONE fused kernel each -- update and both kernel layouts from a single pass
        over p, m, v, g --, every other variable of the slice one multi-range Adam launch, and only the remaining
        layers are repacked separately."""
        V = net.vars
        hi = V.total if hi is None else hi
        fused = getattr(V, "fused_adam_layers", None)
        if not fused:
            self.apply(net, lo, hi, grad_scale)
            net.vars.repack(net.ops, names)
            return
        lr_t = P.adam_lr_t(self.lr, self.t, self.beta1, self.beta2)
        # `fused` is a set of layers (FCN) or {layer: its folded BN gamma variable or None} (GraphNet: the packed weights
        # carry gamma / sqrt(1 + eps), Conv2D_Block of utils.py:186-208).  A folded layer is fused only when its gamma is
        # updated in this very call (before the pack, below); otherwise it takes the separate Adam + repack path, and a
        # gamma updated here whose weights are not packed here gets its layer repacked as well.
        scale_of = fused if isinstance(fused, dict) else {}
        mult = getattr(V, "fold_mult", 1.0)
        here = {name for name, s in V.slots.items() if lo <= s.offset < hi}
        jobs, ranges, rest = [], [], set()
        for name in here:
            layer = name.split("/")[0]
            if name.endswith("/weights") and layer in fused and (scale_of.get(layer) is None or scale_of[layer] in here):
                jobs.append((name, layer))
        packed = {layer for _, layer in jobs}
        owner = {g: layer for layer, g in scale_of.items() if g is not None}
        for name in here:
            layer = name.split("/")[0]
            if name.endswith("/weights") and layer in packed:
                continue
            s = V.slots[name]
            ranges.append((s.offset, s.size))
            if name.endswith("/weights"):
                rest.add(layer)
            elif name in owner and owner[name] not in packed:
                rest.add(owner[name])
        if ranges:
            ranges.sort()
            net.ops.adam_step_ranges(V.p, V.m, V.v, V.g, ranges, lr_t, self.beta1, self.beta2, self.eps, grad_scale)
        for name, layer in jobs:
            gname = scale_of.get(layer)
            net.ops.adam_pack_conv_weights(V.view(V.p, name), V.view(V.m, name), V.view(V.v, name), V.view(V.g, name),
                                           V.wk[layer], V.wd[layer], lr_t, self.beta1, self.beta2, self.eps, grad_scale,
                                           col_scale=V.view(V.p, gname) if gname else None, col_mult=mult if gname else 1.0)
        if rest:
            net.vars.repack(net.ops, rest)


class MomentumOptimizer:
    """tf.train.MomentumOptimizer: a = mu*a + g; p -= lr*a (new functionality)."""

    def __init__(self, learning_rate, momentum=0.9):
        self.lr, self.mu = learning_rate, momentum
        self.t = 0

    def minimize(self, net: FCN, allreduce=None):
        if net.vars.m is None:
            net.vars.m = torch.zeros_like(net.vars.p)
        return TrainStep(net, self, allreduce)

    def apply(self, net: FCN, lo=0, hi=None, grad_scale=1.0):
        V = net.vars
        hi = V.total if hi is None else hi
        net.ops.momentum_step(V.p[lo:hi], V.m[lo:hi], V.g[lo:hi], self.lr, self.mu, grad_scale)


class TrainStep:
    """`train_step` of FCN.py:340; calling it is one `sess.run(train_step, feed_dict)` (FCN.py:398):
    forward, loss, backward, (bucketed gradient all-reduce), optimizer update, weight repack.
    Returns the mean loss as a 0-dim device tensor (no host sync)."""

    def __init__(self, net: FCN, opt, allreduce=None):
        self.net, self.opt, self.allreduce = net, opt, allreduce
        self.local = None
        if allreduce is None and getattr(net, "side", None) is not None and net.side.stream is not None:
            if hasattr(net, "nodes"):
                names = [n.name for n in net.nodes if n.kind in ("conv", "deconv")]
                buckets = P.gradient_buckets_even(net.vars.slots, names)
            else:
                buckets = P.gradient_buckets(net.vars.slots)
            self.local = LocalBuckets(net, opt, buckets)

    def sync_optimizer_state(self):
        """Data parallel with the fused exchange: optimizer slots are sharded over the ranks; call this (on
        every rank) before exporting them, e.g. `checkpoint.save_checkpoint`."""
        if self.allreduce is not None and hasattr(self.allreduce, "gather_optimizer_state"):
            self.allreduce.gather_optimizer_state()

    def _launch_stream(self, net):
        if getattr(self, "_ls", None) is None:
            self._ls = torch.cuda.Stream(net.device)
        return self._ls

    def _finalize_stream(self, net):
        if getattr(self, "_fin", None) is None:
            self._fin = SideStream(net.device, enabled=net.side.stream is not None)
        self._fin.enabled = net.side.enabled
        return self._fin

    def __call__(self, feed_dict=None):
        net, opt = self.net, self.opt
        if feed_dict:
            net.feed(feed_dict)
        net.forward()
        loss = net.loss(with_grad=True)
        opt.t += 1
        if self.local is not None and net.side.enabled:
            # single GPU: per-bucket optimizer update + repack on the side stream, in the shadow of backward
            self.local.begin_step()
            net.backward(after_layer=self.local.layer_done)
            self.local.finish()
        elif self.allreduce is None:
            net.backward()
            net.side.join()
            net.wside.join()
            if hasattr(opt, "apply_and_repack"):
                opt.apply_and_repack(net)
            else:
                opt.apply(net)
                net.vars.repack(net.ops)
        else:
            # data parallel: each bucket's all-reduce is launched as soon as its last layer is done; its
            # optimizer update + repack then run on a second side stream right after the collective,
            # in the shadow of the remaining backward
            self.allreduce.begin_step()
            slots = net.vars.slots
            fin = self._finalize_stream(net)
            # exchange fused with the optimizer (dp.SymmetricAllReduce over NVLS): the collective kernel itself
            # applies Adam to this rank's share and multicasts the new parameters; only the repack is left
            fused = bool(getattr(self.allreduce, "fused", False)) and isinstance(opt, AdamOptimizer)
            if fused:
                self.allreduce.set_adam(net.vars.m, net.vars.v, P.adam_lr_t(opt.lr, opt.t, opt.beta1, opt.beta2),
                                        opt.beta1, opt.beta2, opt.eps)
            elif hasattr(self.allreduce, "_adam"):
                self.allreduce._adam = None

            def finalize(lo, hi, work):
                names = {n.split("/")[0] for n, s in slots.items() if lo <= s.offset < hi}

                def go():
                    if work is not None:
                        work.wait()               # the side stream waits for the collective, not the main one
                    if not fused and hasattr(opt, "apply_and_repack"):
                        opt.apply_and_repack(net, lo, hi, names)
                    else:
                        if not fused:
                            opt.apply(net, lo, hi)
                        net.vars.repack(net.ops, names)

                fin.run(go)

            def layer_done(name):
                if not self.allreduce.fires_at(name):
                    return
                # The bucket's gradients come from three streams (dgrad chain, bias gradients, weight
                # gradients).  The collective is issued from a launch stream that waits for all three, so
                # the main stream never blocks on the wgrad stream and backward keeps its overlap.
                if net.side.enabled:
                    ls = self._launch_stream(net)
                    ls.wait_stream(torch.cuda.current_stream())
                    ls.wait_stream(net.side.stream)
                    if net.wside.enabled:
                        ls.wait_stream(net.wside.stream)
                    with torch.cuda.stream(ls):
                        for lo, hi, work in self.allreduce.layer_done(name):
                            finalize(lo, hi, work)
                else:
                    for lo, hi, work in self.allreduce.layer_done(name):
                        finalize(lo, hi, work)

            net.backward(after_layer=layer_done)
            net.side.join()
            net.wside.join()
            for lo, hi, work in self.allreduce.flush():
                finalize(lo, hi, work)
            fin.join()
        net.step_count += 1
        net._ran_forward = False
        net.mark_step_end()
        return loss
