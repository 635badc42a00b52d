"""conv6 (7x7, 512 -> 4096 on 5x18 maps, B=32) dgrad under different split-K factors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda:0")
ops = Ops(dev)
g = torch.Generator().manual_seed(0)
N, H, W, ci, co, k = 32, 5, 18, 512, 4096, 7


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


x = (torch.randn((N, H, W, ci), generator=g) * 0.5).to(torch.bfloat16).to(dev)
w = (torch.randn((k, k, ci, co), generator=g) * 0.01).to(dev)
wk, wd = ops.pack_conv_weights(w)
y = torch.empty((N, H, W, co), dtype=torch.bfloat16, device=dev)
dy = (torch.randn((N, H, W, co), generator=g) * 0.1).to(torch.bfloat16).to(dev)
dx = torch.empty_like(x)
b = torch.zeros(co, device=dev)
print("fwd", round(timeit(lambda: ops.conv2d_fwd(x, wk, b, y, k, k, relu=True)), 1))
for ks in (0, 1, 2, 3, 4, 5, 6, 8, 11, 14, 16):
    ops.ctx.set_tuning("force_ksplit", ks)
    print("dgrad ks", ks, round(timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask=x)), 1))
ops.ctx.set_tuning("force_ksplit", 0)
