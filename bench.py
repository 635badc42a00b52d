#!/usr/bin/env python
"""Benchmark of the FCN-8s training hot path (BASELINE.json: train images/sec, FCN-8s 160x576).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle restatement)

A "step" is one pass of the hot path over one synthetic batch: forward + mean softmax-xent loss +
backward + TF-Adam update (+ gradient all-reduce when N > 1), FCN.py:398.  Workload at every N:
BASELINE.json configs[1], batch 32 per GPU, 160x576x3 u8 images, bf16 storage / fp32 accumulate.
Prints ONE JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train images/sec FCN-8s 160x576"
UNIT = "images/s"
H, W, CIN, NCLS = 160, 576, 3, 2
BATCH_PER_GPU = 32
TRAIN_GFLOP_PER_IMAGE = 230.78      # valid-tap fwd+dgrad+wgrad, BASELINE.md §4 / SURVEY §8d
KEEP_PROB = 0.8                     # FCN.py:395
SETTLE_S = 1.0                      # idle pause before each timed region (same power / clock state for both)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = upper half (the sampler also sees the idle edges)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_images_per_s(steps, warmup, max_seconds=240.0):
    """The reference's CPU path: fp32 FCN-8s fwd + bwd + TF-Adam at batch 1 (a bounded sample of the
    batch-32 workload) on all host cores.  TensorFlow is not installable here, so this
    is the oracle restatement ("port")."""
    import torch
    from oracle.fcn_oracle import FCN8sOracle, default_threads, init_variables, synthetic_batch
    cores = default_threads()
    torch.set_num_threads(cores)
    orc = FCN8sOracle(init_variables(CIN, NCLS, 4096, seed=1234, init="ref"), threads=cores)
    x, lab = synthetic_batch(1, H, W, CIN, seed=0)
    for _ in range(warmup):
        orc.train_step(x, lab, keep_prob=1.0)
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.train_step(x, lab, keep_prob=1.0)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > max_seconds:
            break
    ms = 1e3 * sum(times) / len(times)
    return 1e3 / ms, ms, cores, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # each step is ~seconds of CPU work: bound the run to a few minutes
    steps, warmup = min(steps, 5), min(warmup, 1)
    ips, ms, cores, done = cpu_reference_images_per_s(steps, warmup)
    sample = f"batch 1 of the 32-image step, fp32, {done} step(s) after {warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "FCN-8s 2-class bf16 training (fwd+loss+bwd+Adam), batch 32 per GPU, 160x576x3 (BASELINE configs[1])",
                   "sample": "CPU reference arm: 1 image of the 32-image step per timed step, fp32, keep_prob 1.0",
                   "global_batch": 1, "keep_prob": 1.0},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (torch-CPU fp32 with TF-1.15 op semantics), not TensorFlow: TF is not installable in this image",
    }
    args.emit(json.dumps(line))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from semanticsegmentation_tensorflow_b200 import build_library
    from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce, SymmetricAllReduce, init_distributed
    from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer
    from semanticsegmentation_tensorflow_b200.ops import Profile

    rank, world, local = init_distributed("nccl")
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch
    steps, warmup = args.steps, max(args.warmup, 3)
    global H, W
    if args.res:
        H, W = (int(v) for v in args.res.lower().split("x"))

    # synthetic KITTI-road-shaped batch (raw 0..255 u8 pixels, class-id labels), per rank
    gen = torch.Generator().manual_seed(1000 + rank)
    host_x = torch.randint(0, 256, (B, H, W, CIN), dtype=torch.uint8, generator=gen).pin_memory()
    host_y = torch.randint(0, 2, (B, H, W), dtype=torch.uint8, generator=gen).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)

    if args.model == "fcdensenet":
        from semanticsegmentation_tensorflow_b200.densenet import FCDenseNet, fcdensenet_flops_per_image
        net = FCDenseNet(dev_x, KEEP_PROB, NCLS, seed=1234, world_size=world, dropout_seed=42 + rank)
        train_gflop = fcdensenet_flops_per_image(net.nodes, net.ch, H, W)[1] / 1e9
        workload = f"FCDenseNet (FCDenseNet.py:83-163, 130 conv layers) 2-class bf16 training (fwd+loss+bwd+Adam), batch {B} per GPU, {H}x{W}x3 (BASELINE configs[4])"
        metric = f"train images/sec FCDenseNet {H}x{W}"
    elif args.model in ("unet", "segnet"):
        from semanticsegmentation_tensorflow_b200.graph import SegNet, UNet, graph_flops_per_image, segnet_nodes, unet_nodes
        build, nodes = (UNet, unet_nodes) if args.model == "unet" else (SegNet, segnet_nodes)
        net = build(dev_x, NCLS, seed=1234, world_size=world)
        train_gflop = graph_flops_per_image(nodes(NCLS), H, W, CIN)[1] / 1e9
        if args.model == "unet":
            workload = "U-Net 2-class bf16 training (fwd+loss+bwd+Adam), batch 32 per GPU, 160x576x3 (BASELINE configs[2])"
            metric = "train images/sec U-Net 160x576"
        else:
            workload = "SegNet (SegNet.py:28-87) 2-class bf16 training (fwd+loss+bwd+Adam), batch 32 per GPU, 160x576x3"
            metric = "train images/sec SegNet 160x576"
    else:
        net = FCN(dev_x, KEEP_PROB, NCLS, init="device", seed=1234, world_size=world, dropout_seed=42 + rank)
        train_gflop = TRAIN_GFLOP_PER_IMAGE
        workload = "FCN-8s 2-class bf16 training (fwd+loss+bwd+Adam), batch 32 per GPU, 160x576x3 (BASELINE configs[1])"
        metric = METRIC
    # gradient exchange: our NVLink kernel on a symmetric arena when the box offers symmetric memory, else NCCL
    allreduce = None
    if world > 1:
        allreduce = SymmetricAllReduce.try_create(net) or BucketedAllReduce.for_net(net)
    exchange = getattr(allreduce, "kind", "nccl") if allreduce is not None else None
    train_step = AdamOptimizer(1e-4).minimize(net, allreduce=allreduce)
    feed_dev = {net.image: dev_x, net.annotation: dev_y, net.keep_probability: KEEP_PROB}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        loss = train_step(feed_dev)
    sync()
    time.sleep(SETTLE_S)
    for _ in range(2):
        loss = train_step(feed_dev)

    # ---- timed region 1: device-resident inputs -> `value` ------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = net.ops.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(steps):
        loss = train_step(feed_dev)
    e1.record()
    sync()
    ms_total = e0.elapsed_time(e1)
    launches = net.ops.ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
    ms_step = ms_total / steps
    value = world * B * steps / (ms_total / 1e3)
    final_loss = float(loss)

    # ---- timed region 2: end to end through the public API with HOST buffers -------------
    # (runs right after region 1, before the profiling pass, behind the same idle pause: both regions start from
    # the same power / clock state -- the boxes run under sw_power_cap, and a region timed after seconds of
    # continuous load sees lower SM clocks than one timed from idle)
    feed_host = {net.image: host_x, net.annotation: host_y, net.keep_probability: KEEP_PROB}
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    for _ in range(2):
        loss_host.copy_(train_step(feed_host).reshape(1), non_blocking=True)
    sync()
    time.sleep(SETTLE_S)
    for _ in range(2):
        loss_host.copy_(train_step(feed_host).reshape(1), non_blocking=True)
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    sync()
    e0.record()
    for _ in range(steps):
        l = train_step(feed_host)                       # H2D of images + labels inside
        loss_host.copy_(l.reshape(1), non_blocking=True)     # D2H of the step's loss
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1)
    clocks_e2e = sampler2.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t[0])
    e2e_value = world * B * steps / (ms_e2e / 1e3)

    # ---- timed region 1b: same K steps with per-launch CUDA events (roofline).  The side streams are
    # disabled here so that every launch's event pair brackets that kernel alone.
    side_was, wside_was = net.side.enabled, net.wside.enabled
    net.side.join()
    net.wside.join()
    net.side.enabled = False
    net.wside.enabled = False
    net.ops.profile = Profile()
    sync()
    e0.record()
    for _ in range(steps):
        loss = train_step(feed_dev)
    e1.record()
    sync()
    ms_serial = e0.elapsed_time(e1) / steps
    prof = net.ops.profile.summary()
    prof_detail = net.ops.profile.summary(detail=True)
    net.ops.profile = None
    net.side.enabled = side_was
    net.wside.enabled = wside_was

    # ---- data-parallel correctness, outside the timed regions (world > 1) ---------------------------
    dp_check = dp_correctness(net, train_step, allreduce, feed_dev, world, dev) if world > 1 else None

    if rank != 0:
        return
    if args.layers_out:
        rows = []
        for k, v in sorted(prof_detail.items(), key=lambda kv: -kv[1]["ms"]):
            rate = v["work"] / (v["ms"] / 1e3) / (1e12 if v["unit"] == "flop" else 1e9) if v["ms"] > 0 else 0.0
            rows.append({"call": k, "ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps,
                         "rate": rate, "rate_unit": "TFLOP/s" if v["unit"] == "flop" else "GB/s"})
        os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
        with open(args.layers_out, "w") as f:
            json.dump(rows, f, indent=1)
    peaks = measured_peaks()
    # dominant kernel = the C entry point (one CUDA kernel template behind it) with the largest share of
    # device time in the step; `traffic` = ncu DRAM bytes per launch averaged over that entry point's
    # launches of one step (profiles/ncu_traffic.json, keyed by call shape), when every shape was captured
    total_prof_ms = sum(v["ms"] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    traffic = None
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            table = json.load(f)
        shapes = {k: v for k, v in prof_detail.items() if k.startswith(dom + "[")}
        if shapes and all(k in table for k in shapes):
            tot_b = sum(table[k]["dram_bytes"] * v["launches"] for k, v in shapes.items())
            traffic = tot_b / sum(v["launches"] for v in shapes.values())
            tp = sum(table[k]["tensor_pipe_active_pct"] * v["ms"] for k, v in shapes.items()) / sum(v["ms"] for v in shapes.values())
            ncu = {"tensor_pipe_active_pct": tp}
    if d["unit"] == "flop":
        achieved = d["work"] / (d["ms"] / 1e3) / 1e12
        peak, unit, bound = peaks["bf16_tflops_sustained"], "TFLOP/s", "tensor"
    else:
        achieved = d["work"] / (d["ms"] / 1e3) / 1e9
        peak, unit, bound = peaks["hbm_gbs"], "GB/s", "hbm"
    burst = peaks["bf16_tflops"] if bound == "tensor" else peaks["hbm_gbs"]
    roofline = {"bound": bound, "kernel": dom, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak, "frac_burst": achieved / burst, "peak_burst": burst,
                "clock_regime": ("timed inside a %.2f s region at SM clock %s MHz (max %s): `frac` is against the sustained "
                                 "(power-capped, ~1340 MHz) cuBLAS figure, `frac_burst` against the burst figure" %
                                 (ms_total / 1e3, clocks.get("sm_mhz") if clocks else None, clocks.get("sm_max_mhz") if clocks else None)),
                "traffic": traffic,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, average over the entry point's launches of one step under ncu (profiles/ncu_traffic.json, r2_ncu_step_tc_launches.md)" if traffic else None,
                "ncu_tensor_pipe_active_pct": ncu.get("tensor_pipe_active_pct"),
                "peak_source": peaks["source"] + (" sustained" if bound == "tensor" else ""),
                "algorithmic_per_launch": d["work"] / d["launches"], "algorithmic_unit": d["unit"],
                "launches": d["launches"], "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / total_prof_ms if total_prof_ms else None,
                "measured_in": "second timed pass of the same K steps with per-launch CUDA events, side stream "
                               "off so each event pair brackets one kernel; that pass ran at %.3f ms/step" % ms_serial}
    conv = [v for k, v in prof.items() if v["unit"] == "flop" and ("conv2d_fwd" in k or "conv2d_dgrad" in k or "conv2d_wgrad" in k) and "small" not in k]
    conv_ms = sum(v["ms"] for v in conv)
    conv_gemm = {"ms_per_step": conv_ms / steps, "tflops": sum(v["work"] for v in conv) / (conv_ms / 1e3) / 1e12 if conv_ms else None,
                 "frac_of_sustained_peak": (sum(v["work"] for v in conv) / (conv_ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"]) if conv_ms else None,
                 "frac_of_burst_peak": (sum(v["work"] for v in conv) / (conv_ms / 1e3) / 1e12 / peaks["bf16_tflops"]) if conv_ms else None}
    families = {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps,
                    ("tflops" if v["unit"] == "flop" else "gbs"):
                        (v["work"] / (v["ms"] / 1e3) / (1e12 if v["unit"] == "flop" else 1e9)) if v["ms"] > 0 else None}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    step_tflops = train_gflop * 1e9 * value / world / 1e12

    secondary = None
    if world == 1 and not args.no_secondary and args.model == "fcn":
        del net, train_step
        torch.cuda.empty_cache()
        secondary = run_secondary(dev, peaks)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        ips, ms, cores, done = cpu_reference_images_per_s(2, 1, max_seconds=60.0)
        cpu_baseline = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"batch 1 of the 32-image step, fp32 oracle, {done} step(s) after 1 warm-up, {ms:.0f} ms/step"}

    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload,
                   "global_batch": world * B, "batch_per_gpu": B, "keep_prob": KEEP_PROB,
                   "parallelism": f"dp{world}", "exchange": exchange, "init": "random N(0,0.01^2) (FCN.py:125)",
                   "l2": "working set (1.8 GiB activations + 2.2 GiB weights/optimizer state) >> 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(host_x.numel() + host_y.numel()), "d2h_bytes_per_step": 4,
                "clocks": clocks_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "tensor_util_step": {"train_tflops_per_gpu": step_tflops, "frac_of_peak": step_tflops / peaks["bf16_tflops_sustained"],
                             "flops_per_image": train_gflop * 1e9, "convention": "valid-tap fwd+dgrad+wgrad"},
        "conv_gemms": conv_gemm,
        "kernel_families": families,
        "cpu_baseline": cpu_baseline,
        "final_loss": final_loss,
    }
    if dp_check is not None:
        line["dp_check"] = dp_check
        line.update({k: dp_check[k] for k in ("replica_max_diff", "fused_vs_nccl_param_rel_err", "grad_checksum",
                                              "grad_checksum_equal_across_ranks") if k in dp_check})
    if secondary is not None:
        line["secondary"] = secondary
    args.emit(json.dumps(line))


def dp_correctness(net, train_step, allreduce, feed, world, dev):
    """One more step from identical state through (a) the exchange the bench ran and (b) plain ncclAllReduce +
    the local Adam kernel; reports whether the replicas stayed bit-identical and how far the two differ.
    Runs on every rank (collectives inside); the timed regions are over."""
    import torch
    import torch.distributed as dist
    from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce
    from semanticsegmentation_tensorflow_b200.fcn import AdamOptimizer
    V, opt = net.vars, train_step.opt

    def spread(t):
        """max over elements of (max over ranks - min over ranks): 0.0 iff bit-identical everywhere."""
        hi, lo = t.clone(), t.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return float((hi - lo).abs().max())

    out = {"exchange": getattr(allreduce, "kind", "nccl")}
    torch.cuda.synchronize()
    out["replica_max_diff_after_timed_steps"] = spread(V.p)
    train_step.sync_optimizer_state()                       # full m / v on every rank (fused path shards them)
    p0, m0, v0 = V.p.clone(), V.m.clone(), V.v.clone()
    t0, sc0 = opt.t, net.step_count

    def restore():
        V.p.copy_(p0); V.m.copy_(m0); V.v.copy_(v0)
        opt.t, net.step_count = t0, sc0
        V.repack(net.ops)
        torch.cuda.synchronize()
        dist.barrier()

    train_step(feed)
    torch.cuda.synchronize()
    p_a = V.p.clone()
    out["replica_max_diff"] = spread(p_a)
    restore()
    nccl_step = AdamOptimizer(opt.lr, opt.beta1, opt.beta2, opt.eps)
    nccl_step.t = t0
    step_b = nccl_step.minimize(net, allreduce=BucketedAllReduce.for_net(net))
    step_b(feed)
    torch.cuda.synchronize()
    p_b = V.p.clone()
    out["replica_max_diff_nccl"] = spread(p_b)
    upd = float((p_b - p0).abs().max())
    out["fused_vs_nccl_param_rel_err"] = float((p_a - p_b).abs().max()) / max(float(p_b.abs().max()), 1e-30)
    out["fused_vs_nccl_update_rel_err"] = float((p_a - p_b).abs().max()) / max(upd, 1e-30)
    # the all-reduced gradient arena must be the same on every rank
    cs = V.g.double().sum().reshape(1)
    allcs = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    vals = [float(c) for c in allcs]
    out["grad_checksum"] = vals[0]
    out["grad_checksum_equal_across_ranks"] = all(v == vals[0] for v in vals)
    out["grad_replica_max_diff"] = spread(V.g)
    restore()
    return out


def run_secondary(dev, peaks):
    """Bounded runs of the other BASELINE configs on the same GPU, after the headline regions: U-Net training
    (configs[2]), FCN-8s 384x1248 batch-16 inference (configs[3]), FCDenseNet training (the model of configs[4]) and
    batch-1 160x576 inference latency."""
    import torch
    from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer
    from semanticsegmentation_tensorflow_b200.graph import UNet, graph_flops_per_image, unet_nodes
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gen = torch.Generator().manual_seed(2000)

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # U-Net training, batch 32, host buffers in (e2e), loss out
    B = BATCH_PER_GPU
    hx = torch.randint(0, 256, (B, H, W, CIN), dtype=torch.uint8, generator=gen).pin_memory()
    hy = torch.randint(0, 2, (B, H, W), dtype=torch.uint8, generator=gen).pin_memory()
    net = UNet(hx.to(dev), NCLS, seed=1234)
    step = AdamOptimizer(1e-4).minimize(net)
    feed = {net.image: hx, net.annotation: hy}
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    ms = timed(lambda: loss_host.copy_(step(feed).reshape(1), non_blocking=True), 10)
    gf = graph_flops_per_image(unet_nodes(NCLS), H, W, CIN)[1]
    tf = gf * B / (ms / 1e3) / 1e12
    out["unet_train_160x576_b32"] = {"images_per_s": B / (ms / 1e3), "ms_per_step": ms, "tflops": tf,
                                     "frac_of_burst_peak": tf / peaks["bf16_tflops"], "steps": 10,
                                     "workload": "BASELINE configs[2] at 1 GPU, host buffers in / loss out"}
    del net, step
    torch.cuda.empty_cache()
    # the reference's SegNet (SegNet.py:28-87), same workload
    from semanticsegmentation_tensorflow_b200.graph import SegNet, segnet_nodes
    net = SegNet(hx.to(dev), NCLS, seed=1234)
    step = AdamOptimizer(1e-4).minimize(net)
    feed = {net.image: hx, net.annotation: hy}
    ms = timed(lambda: loss_host.copy_(step(feed).reshape(1), non_blocking=True), 10)
    gf = graph_flops_per_image(segnet_nodes(NCLS), H, W, CIN)[1]
    tf = gf * B / (ms / 1e3) / 1e12
    out["segnet_train_160x576_b32"] = {"images_per_s": B / (ms / 1e3), "ms_per_step": ms, "tflops": tf,
                                       "frac_of_burst_peak": tf / peaks["bf16_tflops"], "steps": 10,
                                       "workload": "the reference's SegNet (SegNet.py:28-87) at 1 GPU, host buffers in / loss out"}
    del net, step
    torch.cuda.empty_cache()
    # FCN-8s full-resolution inference, batch 16, host images in / road masks out
    h2, w2, B2 = 384, 1248, 16
    hx2 = torch.randint(0, 256, (B2, h2, w2, CIN), dtype=torch.uint8, generator=gen).pin_memory()
    mask_host = torch.empty((B2, h2, w2), dtype=torch.uint8).pin_memory()
    net = FCN(hx2.to(dev), 1.0, NCLS, init="device", seed=1234)
    ms = timed(lambda: mask_host.copy_(net.infer(hx2)[1], non_blocking=True), 20)
    tf = 428.29e9 * B2 / (ms / 1e3) / 1e12
    out["fcn_infer_384x1248_b16"] = {"images_per_s": B2 / (ms / 1e3), "ms_per_batch": ms, "tflops": tf,
                                     "frac_of_burst_peak": tf / peaks["bf16_tflops"], "iters": 20,
                                     "workload": "BASELINE configs[3], host images in / road masks out"}
    del net
    torch.cuda.empty_cache()
    # FCDenseNet (FCDenseNet.py:83-163, BASELINE configs[4]'s model) training, batch 32, host buffers in, loss out
    from semanticsegmentation_tensorflow_b200.densenet import FCDenseNet, fcdensenet_flops_per_image
    net = FCDenseNet(hx.to(dev), KEEP_PROB, NCLS, seed=1234)
    step = AdamOptimizer(1e-4).minimize(net)
    feed = {net.image: hx, net.annotation: hy, net.keep_probability: KEEP_PROB}
    ms = timed(lambda: loss_host.copy_(step(feed).reshape(1), non_blocking=True), 6)
    gf = fcdensenet_flops_per_image(net.nodes, net.ch, H, W)[1]
    tf = gf * B / (ms / 1e3) / 1e12
    out["fcdensenet_train_160x576_b32"] = {"images_per_s": B / (ms / 1e3), "ms_per_step": ms, "tflops": tf,
                                           "frac_of_burst_peak": tf / peaks["bf16_tflops"], "steps": 6,
                                           "workload": "FCDenseNet (103-layer Tiramisu of FCDenseNet.py, BASELINE configs[4]'s model) at "
                                                       "160x576, 1 GPU, keep_prob 0.8, host buffers in / loss out; logical-channel FLOPs"}
    del net, step
    torch.cuda.empty_cache()
    # batch-1 latency at 160x576 (gen_test_output's per-image call, FCN.py:224-231)
    hx1 = hx[:1].clone().pin_memory()
    mask1 = torch.empty((1, H, W), dtype=torch.uint8).pin_memory()
    net = FCN(hx1.to(dev), 1.0, NCLS, init="device", seed=1234)
    lat = []
    for i in range(53):
        e0.record()
        mask1.copy_(net.infer(hx1)[1], non_blocking=True)
        e1.record()
        e1.synchronize()
        if i >= 3:
            lat.append(e0.elapsed_time(e1))
    lat.sort()
    out["fcn_infer_160x576_b1_latency_ms"] = {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(0.99 * len(lat)))],
                                              "min": lat[0], "iters": len(lat)}
    # the same call replayed from a CUDA graph (net.infer_graphed: one H2D copy + one graph launch + the D2H of the mask)
    glat = []
    for i in range(53):
        e0.record()
        mask1.copy_(net.infer_graphed(hx1)[1], non_blocking=True)
        e1.record()
        e1.synchronize()
        if i >= 3:
            glat.append(e0.elapsed_time(e1))
    glat.sort()
    out["fcn_infer_160x576_b1_latency_ms"]["cuda_graph"] = {"p50": glat[len(glat) // 2],
                                                            "p99": glat[min(len(glat) - 1, int(0.99 * len(glat)))], "min": glat[0]}
    return out


def run_infer(args):
    """BASELINE configs[3]: full-resolution inference, batch 16, one GPU; host images in, road mask out."""
    import torch
    from semanticsegmentation_tensorflow_b200 import build_library
    from semanticsegmentation_tensorflow_b200.fcn import FCN
    build_library()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    h, w, B = 384, 1248, 16
    gen = torch.Generator().manual_seed(7)
    host_x = torch.randint(0, 256, (B, h, w, CIN), dtype=torch.uint8, generator=gen).pin_memory()
    net = FCN(host_x.to(dev), 1.0, NCLS, init="device", seed=1234)
    mask_host = torch.empty((B, h, w), dtype=torch.uint8).pin_memory()
    for _ in range(max(args.warmup, 3)):
        net.infer(host_x)
    torch.cuda.synchronize()
    lat = []
    launches0 = net.ops.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for _ in range(args.steps):
        e0.record()
        prob, mask = net.infer(host_x)              # H2D of 16 images inside
        mask_host.copy_(mask, non_blocking=True)    # D2H of the road masks
        e1.record()
        e1.synchronize()
        lat.append(e0.elapsed_time(e1))
    t_all1.record()
    torch.cuda.synchronize()
    total_ms = t_all0.elapsed_time(t_all1)
    lat.sort()
    fwd_gflop = 428.29
    line = {
        "metric": "inference images/sec FCN-8s 384x1248", "value": B * args.steps / (total_ms / 1e3), "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "FCN-8s 2-class forward + softmax + road mask, batch 16, 384x1248x3 (BASELINE configs[3])",
                   "global_batch": B, "keep_prob": 1.0},
        "latency_ms": {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(0.99 * len(lat)))], "min": lat[0]},
        "e2e": {"value": B * args.steps / (total_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(host_x.numel()), "d2h_bytes_per_step": int(mask_host.numel())},
        "gpu_launches": int(net.ops.ctx.launches - launches0),
        "tensor_util_step": {"fwd_tflops": fwd_gflop * 1e9 * B * args.steps / (total_ms / 1e3) / 1e12,
                             "flops_per_image": fwd_gflop * 1e9, "convention": "valid-tap forward"},
    }
    args.emit(json.dumps(line))


class _StdoutGuard:
    """Everything except the final JSON line goes to stderr (NCCL / torchrun print banners on fd 1)."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU per step")
    ap.add_argument("--res", default=None, help="HxW of the synthetic images for the secondary models (default 160x576), e.g. 384x1248")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bounded U-Net / inference runs after the headline regions")
    ap.add_argument("--layers-out", default=None, help="write the per-call (per-layer) timing table here")
    ap.add_argument("--model", default="fcn", choices=["fcn", "unet", "segnet", "fcdensenet"],
                    help="fcn = FCN-8s (BASELINE configs[1], the driver's metric); unet = configs[2]; segnet = the reference's SegNet")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train = BASELINE configs[1] (default, the driver's metric); infer = configs[3]: "
                         "FCN-8s forward + softmax + road mask at 384x1248, batch 16 (throughput and latency)")
    args = ap.parse_args()
    with _StdoutGuard() as out:
        args.emit = out.emit
        if args.impl == "reference":
            run_reference(args)
        elif args.workload == "infer":
            run_infer(args)
        else:
            run_gpu(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
