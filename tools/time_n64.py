"""CUDA-event timings of the full-resolution 64/128-channel layers (B=32, 160x576 / 80x288) under
the kernel-selection switches (slab3 on/off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda:0")
ops = Ops(dev)
g = torch.Generator().manual_seed(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def case(N, H, W, ci, co):
    x = (torch.randn((N, H, W, ci), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn((3, 3, ci, co), generator=g) * 0.01).to(dev)
    b = torch.zeros(co, device=dev)
    wk, wd = ops.pack_conv_weights(w)
    y = torch.empty((N, H, W, co), dtype=torch.bfloat16, device=dev)
    dy = (torch.randn((N, H, W, co), generator=g) * 0.1).to(torch.bfloat16).to(dev)
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    out = {}
    for s3 in (2, 1, 0):
        ops.ctx.set_tuning("slab3", s3)
        out[f"fwd s3={s3}"] = timeit(lambda: ops.conv2d_fwd(x, wk, b, y, 3, 3, relu=True))
        out[f"dgrad s3={s3}"] = timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, 3, 3, relu_mask=x))
    ops.ctx.set_tuning("slab3", 1)
    out["wgrad"] = timeit(lambda: ops.conv2d_wgrad(x, dy, dw, 3, 3))
    print((N, H, W, ci, co), {k: round(v, 1) for k, v in out.items()})


case(32, 160, 576, 128, 64)
case(32, 160, 576, 64, 64)
case(32, 80, 288, 64, 128)
case(32, 80, 288, 128, 128)
case(16, 384, 1248, 64, 64)
