"""conv6 (7x7, 512 -> 4096 on 5x18 maps, B=32): forward, dgrad under the team stream-K / plain split-K schedules,
wgrad.  Plain run prints CUDA-event times; under ncu (`-k regex:'^(igemm_kernel|wgrad_kernel)'`) the launch order is
the printed one (PROBE_ONCE=1: one launch each, no timing loops)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda:0")
ops = Ops(dev)
g = torch.Generator().manual_seed(0)
N, H, W, ci, co, k = 32, 5, 18, 512, 4096, 7
once = os.environ.get("PROBE_ONCE") == "1"


def timeit(fn, iters=10):
    if once:
        fn()
        torch.cuda.synchronize()
        return 0.0
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


x = (torch.randn((N, H, W, ci), generator=g) * 0.5).clamp_min(0).to(torch.bfloat16).to(dev)
w = (torch.randn((k, k, ci, co), generator=g) * 0.01).to(dev)
wk, wd = ops.pack_conv_weights(w)
y = torch.empty((N, H, W, co), dtype=torch.bfloat16, device=dev)
dy = (torch.randn((N, H, W, co), generator=g) * 0.1).to(torch.bfloat16).to(dev)
dx = torch.empty_like(x)
dw = torch.empty_like(w)
b = torch.zeros(co, device=dev)
print("fwd", round(timeit(lambda: ops.conv2d_fwd(x, wk, b, y, k, k, relu=True)), 1))
for mode in os.environ.get("PROBE_MODES", "team,ks2,ks3,ks6").split(","):
    if mode == "team":
        ops.ctx.set_tuning("teamk", 1); ops.ctx.set_tuning("force_ksplit", 0)
    else:
        ops.ctx.set_tuning("teamk", 0); ops.ctx.set_tuning("force_ksplit", int(mode[2:]))
    print("dgrad", mode, round(timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask=x)), 1))
ops.ctx.set_tuning("teamk", 1); ops.ctx.set_tuning("force_ksplit", 0)
print("wgrad", round(timeit(lambda: ops.conv2d_wgrad(x, dy, dw, k, k)), 1))
# conv7 for reference
x7 = (torch.randn((N, H, W, co), generator=g) * 0.5).clamp_min(0).to(torch.bfloat16).to(dev)
w7 = (torch.randn((1, 1, co, co), generator=g) * 0.01).to(dev)
wk7, wd7 = ops.pack_conv_weights(w7)
dw7 = torch.empty_like(w7)
print("conv7 fwd", round(timeit(lambda: ops.conv2d_fwd(x7, wk7, b, y, 1, 1, relu=True)), 1))
print("conv7 dgrad", round(timeit(lambda: ops.conv2d_dgrad(dy, wd7, x7.clone(), 1, 1)), 1))
print("conv7 wgrad", round(timeit(lambda: ops.conv2d_wgrad(x7, dy, dw7, 1, 1)), 1))
