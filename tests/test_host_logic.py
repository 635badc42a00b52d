"""CPU tests: host-side planning logic, the C-ABI library (loads, exports every symbol the
header declares), and the data-parallel bucket logic over gloo with world_size 2."""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from semanticsegmentation_tensorflow_b200 import plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_table_matches_reference_graph():
    L = P.fcn8s_layers(3, 2)
    convs = [l for l in L if l.kind == "conv"]
    assert len(convs) == 17 and sum(l.kind == "pool" for l in L) == 5 and sum(l.kind == "deconv" for l in L) == 3
    assert [l.k for l in convs].count(3) == 14                       # FCN.py:52-73
    assert convs[-3].k == 7 and convs[-3].cout == 4096 and convs[-3].dropout      # conv6, FCN.py:78-79
    assert convs[-1].cout == 2 and convs[-1].relu                    # conv8 ReLU'd, FCN.py:86
    t = [l for l in L if l.kind == "deconv"]
    assert [(l.k, l.stride, l.cin, l.cout) for l in t] == [(4, 2, 2, 512), (4, 2, 512, 256), (16, 8, 256, 2)]
    assert t[2].bias_name == "bias"                                  # FCN.py:103


def test_variable_shapes_match_oracle():
    from oracle.fcn_oracle import variable_shapes
    for cin in (3, 4):
        assert list(P.variable_shapes(cin, 2).items()) == list(variable_shapes(cin, 2).items())
    assert sum(int(np.prod(s)) for s in P.variable_shapes().values()) == 138_873_924


def test_reference_init_matches_oracle_stream():
    from oracle.fcn_oracle import init_variables
    from semanticsegmentation_tensorflow_b200.fcn import reference_init
    for init in ("ref", "he"):
        a = reference_init(P.variable_shapes(3, 2, 64), 1234, init)
        b = init_variables(3, 2, 64, 1234, init)
        assert list(a) == list(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_arena_layout_alignment_and_buckets():
    slots, total = P.arena_layout(P.variable_shapes())
    off = 0
    for s in slots.values():
        assert s.offset % P.ALIGN == 0 and s.offset >= off
        off = s.offset + s.size
    assert total >= off and total % P.ALIGN == 0
    b = P.gradient_buckets(slots)
    # backward completion order; contiguous, disjoint, covering
    assert [x[2] for x in b] == ["conv7", "conv6", "conv4_1", "conv1_1"]
    assert b[0][1] == total and b[-1][0] == 0
    for (lo, hi, _), (lo2, hi2, _) in zip(b, b[1:]):
        assert lo == hi2 and lo2 < hi2
    conv6 = slots["conv6/weights"]
    assert b[1][0] == conv6.offset and b[1][1] - b[1][0] >= conv6.size


def test_adam_lr_t_and_sharding():
    assert P.adam_lr_t(1e-4, 1) == pytest.approx(1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9))
    assert P.shard_batch(64, 8, 3) == (24, 32)
    with pytest.raises(ValueError):
        P.shard_batch(10, 4, 0)


def test_flop_model_matches_survey():
    fwd_d, train_d = P.train_flops_per_image(160, 576, valid_taps=False)
    assert fwd_d / 1e9 == pytest.approx(86.58, abs=0.01)
    assert train_d / 1e9 == pytest.approx(259.42, abs=0.01)
    fwd_v, _ = P.train_flops_per_image(160, 576, valid_taps=True)
    assert fwd_v / 1e9 == pytest.approx(77.0, abs=0.2)


def test_library_builds_loads_and_exports_every_declared_symbol():
    from semanticsegmentation_tensorflow_b200 import build_library
    from semanticsegmentation_tensorflow_b200 import _lib
    path = build_library()
    decls = _lib.parse_header()
    assert len(decls) >= 30
    cdll = ctypes.CDLL(path)
    for name in decls:
        assert hasattr(cdll, name), f"libsegk.so does not export {name}"
    assert cdll.segk_abi_version() == 1
    # no GPU here: creating a context must fail with a status, not crash, and nothing falls back
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        cdll.segk_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        assert cdll.segk_create(0, ctypes.byref(h)) != 0 and not h.value
        from semanticsegmentation_tensorflow_b200.fcn import FCN
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            FCN(torch.zeros((1, 32, 32, 3), dtype=torch.uint8), 1.0, 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semanticsegmentation_tensorflow_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from semanticsegmentation_tensorflow_b200 import plan as P
from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
slots, total = P.arena_layout(P.variable_shapes(3, 2, 64))
torch.manual_seed(0)
full = torch.randn(world, total)                      # per-rank gradient contributions
g = full[rank].clone()
ar = BucketedAllReduce(g, P.gradient_buckets(slots))
order = []
ar.begin_step()
layers = [l for l in P.fcn8s_layers(3, 2, 64) if l.kind != "pool"]
for l in reversed(layers):                           # backward visits layers in reverse
    before = len(ar._works)
    ar.layer_done(l.name)
    if len(ar._works) > before:
        order.append(l.name)
done = list(ar.finish())
assert order == ["conv7", "conv6", "conv4_1", "conv1_1"], order
assert sorted(done) == sorted((lo, hi) for lo, hi, _ in ar.buckets)
assert torch.allclose(g, full.sum(0), atol=1e-5), "all-reduce result != sum over ranks"
lo, hi = P.shard_batch(8, world, rank)
assert hi - lo == 8 // world
dist.barrier()
print("rank", rank, "ok")
'''


def test_bucketed_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_generic_even_buckets_cover_the_arena():
    from semanticsegmentation_tensorflow_b200.graph import graph_variable_shapes, unet_nodes
    nodes = unet_nodes(2)
    slots, total = P.arena_layout(graph_variable_shapes(nodes, 3))
    names = [n.name for n in nodes if n.kind in ("conv", "deconv")]
    b = P.gradient_buckets_even(slots, names, 4)
    assert b[0][1] == total and b[-1][0] == 0 and 2 <= len(b) <= 5
    for (lo, hi, _), (lo2, hi2, _) in zip(b, b[1:]):
        assert lo == hi2 and lo2 < hi2
    # backward completion order: the completing layers appear in reverse creation order
    idx = [names.index(x[2]) for x in b]
    assert idx == sorted(idx, reverse=True)


def test_exchange_shares_partition_every_bucket():
    """dp.SymmetricAllReduce.share() must be the ownership rule of csrc/exchange.cu: rank r owns float4 units
    [off4 + r*ceil(n4/world), ...) of a bucket -- a partition of the bucket for every world size."""
    from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce, SymmetricAllReduce
    slots, total = P.arena_layout(P.variable_shapes(3, 2, 4096))
    buckets = P.gradient_buckets(slots)
    by_offset = sorted((lo, hi) for lo, hi, _ in buckets)          # (listed in backward-completion order)
    assert by_offset[0][0] == 0 and by_offset[-1][1] == total
    assert all(by_offset[i][1] == by_offset[i + 1][0] for i in range(len(by_offset) - 1))
    for world in (2, 3, 4, 8):
        for lo, hi, _ in buckets:
            assert lo % 4 == 0 and (hi - lo) % 4 == 0          # float4 granularity of the kernels
            cover = []
            for r in range(world):
                o = object.__new__(SymmetricAllReduce)
                o.world, o.rank = world, r
                a, b = o.share(lo, hi)
                assert lo <= a <= b <= hi and (a - lo) % 4 == 0
                cover.append((a, b))
            assert cover[0][0] == lo and cover[-1][1] == hi
            assert all(cover[i][1] == cover[i + 1][0] or cover[i + 1][0] == cover[i + 1][1] for i in range(world - 1))
    ar = BucketedAllReduce(torch.zeros(total), buckets)
    assert ar.fires_at(buckets[0][2]) and not ar.fires_at("conv1_1")


class _FakeVars:
    """CPU stand-in for fcn.Variables: enough for checkpoint.state_dict / load_state_dict."""

    def __init__(self, shapes):
        self.slots, self.total = P.arena_layout(shapes)
        self.p = torch.arange(self.total, dtype=torch.float32) * 1e-3
        self.m = torch.ones(self.total) * 0.5
        self.v = torch.ones(self.total) * 0.25
        self.repacked = 0

    def view(self, arena, name):
        s = self.slots[name]
        return arena[s.offset:s.offset + s.size].view(s.shape)

    def repack(self, ops):
        self.repacked += 1


class _FakeNet:
    def __init__(self):
        self.vars = _FakeVars(P.variable_shapes(3, 2, 64))
        self.ops = None


@pytest.mark.parametrize("t", [0, 3, 980, 986, 2000, 123456])
def test_checkpoint_step_count_roundtrip_beyond_float32_underflow(tmp_path, t):
    """0.9^(t+1) underflows float32 near t = 986: the step count travels as an int64 `global_step`."""
    from semanticsegmentation_tensorflow_b200.checkpoint import load_checkpoint, save_checkpoint, state_dict
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer, MomentumOptimizer
    net, opt = _FakeNet(), AdamOptimizer(1e-4)
    opt.t = t
    sd = state_dict(net, opt)
    assert int(sd["global_step"]) == t and sd["beta1_power"].dtype == np.float32
    assert "conv6/weights/Adam" in sd and "conv6/weights/Adam_1" in sd and "conv_t3/bias/Adam" in sd
    path = str(tmp_path / "c.npz")
    save_checkpoint(path, net, opt)
    net2, opt2 = _FakeNet(), AdamOptimizer(1e-4)
    net2.vars.p.zero_(); net2.vars.m.zero_(); net2.vars.v.zero_()
    load_checkpoint(path, net2, opt2)
    assert opt2.t == t and net2.vars.repacked == 1
    for name in net.vars.slots:      # (the alignment padding between variables is not part of a checkpoint)
        for arena in ("p", "m", "v"):
            assert torch.equal(net2.vars.view(getattr(net2.vars, arena), name), net.vars.view(getattr(net.vars, arena), name))
    # checkpoints written before global_step existed: beta1_power is inverted only while float32 resolves it
    old = {k: v for k, v in sd.items() if k != "global_step"}
    from semanticsegmentation_tensorflow_b200.checkpoint import load_state_dict
    opt3 = AdamOptimizer(1e-4)
    if t < 900:
        load_state_dict(_FakeNet(), old, opt3)
        assert opt3.t == t
    elif t >= 2000:
        with pytest.raises(ValueError, match="underflowed"):
            load_state_dict(_FakeNet(), old, opt3)
    # tf.train.MomentumOptimizer names its slot <var>/Momentum
    mopt = MomentumOptimizer(1e-3)
    mopt.t = t
    net.vars.v = None
    sdm = state_dict(net, mopt)
    assert "conv6/weights/Momentum" in sdm and "conv6/weights/Adam" not in sdm and "beta1_power" not in sdm
    net4, mopt2 = _FakeNet(), MomentumOptimizer(1e-3)
    net4.vars.m.zero_()
    load_state_dict(net4, sdm, mopt2)
    assert mopt2.t == t and all(torch.equal(net4.vars.view(net4.vars.m, n), net.vars.view(net.vars.m, n)) for n in net.vars.slots)


def test_sharded_optimizer_slots_are_gathered_before_export():
    """With the fused data-parallel exchange every rank holds 1/world of Adam's m / v: state_dict must gather
    first -- through the TrainStep when given, else through the exchange attached to the net."""
    from semanticsegmentation_tensorflow_b200.checkpoint import state_dict
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer

    class Ex:
        fused, slots_stale, calls = True, True, 0

        def gather_optimizer_state(self):
            self.calls += 1
            self.slots_stale = False

    net = _FakeNet()
    net.exchange = Ex()
    state_dict(net, AdamOptimizer(1e-4))
    assert net.exchange.calls == 1 and not net.exchange.slots_stale
    state_dict(net, AdamOptimizer(1e-4))
    assert net.exchange.calls == 1          # nothing stale: no second collective


def test_reference_helper_signatures():
    """The drop-in helpers keep the reference's argument names and defaults (FCN.py:117,138,161,165,169)."""
    import inspect
    from semanticsegmentation_tensorflow_b200 import layers
    def sig(f):
        return [(n, p.default) for n, p in inspect.signature(f).parameters.items() if n != "store" and n != "seed"]
    E = inspect.Parameter.empty
    assert sig(layers.conv_layer) == [("x", E), ("num_filters", E), ("name", E), ("filter_height", 3), ("filter_width", 3),
                                      ("stride", 1), ("padding", "SAME")]
    assert sig(layers.deconv_layer) == [("x", E), ("shape", E), ("num_filters", E), ("name", E), ("output_shape", E),
                                        ("filter_height", 4), ("filter_width", 4), ("stride", 2), ("padding", "SAME")]
    assert sig(layers.max_pool) == [("x", E), ("name", E), ("filter_height", 2), ("filter_width", 2), ("stride", 2),
                                    ("padding", "VALID")]
    assert sig(layers.dropout) == [("x", E), ("keep_prob", E)]
    assert sig(layers.fuse) == [("x1", E), ("x2", E), ("name", E)]
    import semanticsegmentation_tensorflow_b200 as pkg
    assert pkg.conv_layer is layers.conv_layer and pkg.fuse is layers.fuse


def _plan_shapes(nodes, cin=3, hw=(64, 64)):
    """Shapes / routes of a node list as GraphNet._plan derives them (tensor-core route = both channel counts % 64 == 0)."""
    shape, route = {"input": (2, hw[0], hw[1], cin)}, {}
    for n in nodes:
        N, h, w, c = shape[n.inputs[0]]
        if n.kind == "pool":
            shape[n.name] = (N, h // 2, w // 2, c)
        elif n.kind == "concat":
            shape[n.name] = (N, h, w, sum(shape[i][3] for i in n.inputs))
        elif n.kind == "deconv":
            shape[n.name], route[n.name] = (N, h * 2, w * 2, n.cout), "tc"
        else:
            shape[n.name] = (N, h, w, n.cout)
            route[n.name] = "tc" if (c % 64 == 0 and n.cout % 64 == 0) else "other"
    return shape, route


def test_zero_copy_concat_planning():
    """Which Concat inputs (utils.py:332) become channel slices of the concat buffer (graph.concat_slots)."""
    from semanticsegmentation_tensorflow_b200.graph import GraphBuilder, concat_slots, unet_nodes
    nodes = unet_nodes(2)
    shape, route = _plan_shapes(nodes)
    slots = concat_slots(nodes, shape, route)
    # the U-Net: every upsampled half at offset 0, every encoder skip behind it
    assert slots == {"unpool1": ("concat1", 0), "conv13": ("concat1", 512), "unpool2": ("concat2", 0), "conv10": ("concat2", 512),
                     "unpool3": ("concat3", 0), "conv7": ("concat3", 256), "unpool4": ("concat4", 0), "conv4": ("concat4", 128),
                     "unpool5": ("concat5", 0), "conv2": ("concat5", 64)}
    # a ReLU conv read by the Concat and by another conv: its ReluGrad cannot be applied in place -> stays a copy;
    # a conv without ReLU read by the Concat alone, and a pool-only skip whose pool comes AFTER the concat -> copy for the latter
    g = GraphBuilder()
    a = g.Conv2D_Block("input", 64, batch_normalization=True, relu=True, name="a")        # from 3 channels: not a tensor-core route
    b = g.Conv2D_Block(a, 64, batch_normalization=True, relu=True, name="b")             # read by c and the concat
    c = g.Conv2D_Block(b, 64, batch_normalization=True, relu=False, name="c")            # no ReLU, read by the concat only
    d = g.Conv2D_Block(b, 64, batch_normalization=True, relu=True, name="d")             # ReLU, concat + a LATER pool
    cat = g.Concat([a, b, c, d], "cat")
    g.Max_Pooling(d, "late_pool")
    g.Conv2D_Block(cat, 2, 1, 1, name="head")
    shape, route = _plan_shapes(g.nodes)
    assert concat_slots(g.nodes, shape, route) == {"c": ("cat", 128)}
    # the same tensor twice in one Concat, and a concat that is the network output, never alias
    g = GraphBuilder()
    a = g.Conv2D_Block("input", 64, name="a")
    b = g.Conv2D_Block(a, 64, name="b")
    g.Concat([b, b], "cat")
    shape, route = _plan_shapes(g.nodes)
    assert concat_slots(g.nodes, shape, route) == {}


def test_channel_slice_view_pitch():
    """ops._pitch: the channel pitch handed to segk_set_pitch for a torch view, and what is rejected."""
    import torch
    from semanticsegmentation_tensorflow_b200.ops import _pitch
    wide = torch.zeros((2, 4, 6, 256), dtype=torch.bfloat16)
    assert _pitch(wide) == 0 and _pitch(None) == 0
    assert _pitch(wide[..., 64:192]) == 256 and _pitch(wide[..., :64]) == 256
    with pytest.raises(ValueError):
        _pitch(wide.permute(0, 2, 1, 3))                     # not a channel slice
    with pytest.raises(ValueError):
        _pitch(wide[..., 4:68])                              # first element not 16-byte aligned
    assert _pitch(wide[:, :, ::2, :64]) == 512              # every other pixel of a row is still a uniform pixel pitch
    with pytest.raises(ValueError):
        _pitch(wide[:, ::2, :, :64])                         # rows skipped: not one pitch
    odd = torch.zeros((1, 2, 2, 100), dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        _pitch(odd[..., :64])                                # pitch not a multiple of 8 channels


def test_adam_apply_and_repack_planning_for_folded_bn_layers():
    """AdamOptimizer.apply_and_repack on a graph model (BN scale folded into the packed weights): which variables take the
    fused Adam + fold + repack kernel, in which order, and what happens when a gradient bucket separates a layer's weights
    from its gamma.  Pure host logic: the ops are recorded, nothing runs."""
    import torch
    from collections import OrderedDict
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    from semanticsegmentation_tensorflow_b200.graph import _Vars, BN_SCALE

    shapes = OrderedDict([("conv1/weights", (3, 3, 3, 64)), ("bn0/gamma", (64,)), ("bn0/beta", (64,)),
                          ("conv2/weights", (3, 3, 64, 64)), ("bn1/gamma", (64,)), ("bn1/beta", (64,)),
                          ("up/weights", (4, 4, 64, 64)),
                          ("conv3/weights", (3, 3, 64, 128)), ("bn2/gamma", (128,)), ("bn2/beta", (128,)),
                          ("head/weights", (1, 1, 128, 2))])
    V = _Vars(shapes, "cpu", {})
    V.m, V.v = torch.zeros_like(V.p), torch.zeros_like(V.p)
    V.fold_mult = BN_SCALE
    V.fused_adam_layers = {"conv2": "bn1/gamma", "conv3": "bn2/gamma"}      # tensor-core routes; conv1 (first layer), up, head are not
    V.wk = {"conv2": "wk2", "conv3": "wk3"}
    V.wd = {"conv2": "wd2", "conv3": "wd3"}
    calls = []

    class Rec:
        def adam_step_ranges(self, p, m, v, g, ranges, *a):
            calls.append(("ranges", list(ranges)))

        def adam_pack_conv_weights(self, p, m, v, g, wk, wd, *a, col_scale=None, col_mult=1.0):
            calls.append(("pack", wk, None if col_scale is None else col_scale.data_ptr(), col_mult))

    V._repack = lambda ops, only: calls.append(("repack", set(only)))

    class Net:
        vars, ops = V, Rec()

    opt = AdamOptimizer(1e-4)
    opt.t = 1                      # (TrainStep increments the step count before the update)
    off = {k: (s.offset, s.size) for k, s in V.slots.items()}

    # the whole arena: small variables first (gamma is updated before its weights are packed with it), then the fused layers
    opt.apply_and_repack(Net)
    assert calls[0][0] == "ranges"
    assert sorted(calls[0][1]) == sorted(off[k] for k in shapes if k not in ("conv2/weights", "conv3/weights"))
    packs = [c for c in calls if c[0] == "pack"]
    assert {c[1] for c in packs} == {"wk2", "wk3"} and all(c[3] == BN_SCALE for c in packs)
    assert {c[2] for c in packs} == {V.view(V.p, "bn1/gamma").data_ptr(), V.view(V.p, "bn2/gamma").data_ptr()}
    assert calls.index(packs[0]) > 0 and calls[-1] == ("repack", {"conv1", "up", "head"})

    # a bucket that ends between conv2's weights and its gamma: the layer takes the separate path and is repacked now...
    calls.clear()
    lo, hi = off["conv1/weights"][0], off["bn1/gamma"][0]
    opt.apply_and_repack(Net, lo, hi)
    assert [c[0] for c in calls] == ["ranges", "repack"]
    assert off["conv2/weights"] in calls[0][1] and calls[1][1] == {"conv1", "conv2"}
    # ... and again when the bucket with its gamma is applied (the packed copy must carry the new gamma)
    calls.clear()
    opt.apply_and_repack(Net, off["bn1/gamma"][0], off["conv3/weights"][0])
    assert [c[0] for c in calls] == ["ranges", "repack"] and calls[1][1] == {"conv2", "up"}
    # a bucket holding conv3's weights AND gamma fuses it
    calls.clear()
    opt.apply_and_repack(Net, off["conv3/weights"][0], V.total)
    assert [c[0] for c in calls] == ["ranges", "pack", "repack"] and calls[1][1] == "wk3" and calls[2][1] == {"head"}


def test_measurement_scripts_compile():
    """bench.py, __graft_entry__.py and every script under tools/ at least parse (they only run on a GPU box)."""
    import glob
    import py_compile
    for path in [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))):
        py_compile.compile(path, doraise=True)
