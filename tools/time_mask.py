"""dgrad with and without the ReLU mask of the producer layer (B=32 FCN-8s shapes): the time the mask read costs."""
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


for (N, H, W, ci, co) in [(32, 160, 576, 64, 64), (32, 80, 288, 128, 128), (32, 40, 144, 256, 256), (32, 20, 72, 512, 512), (32, 10, 36, 512, 512)]:
    x = (torch.randn((N, H, W, ci), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn((3, 3, ci, co), generator=g) * 0.01).to(dev)
    wk, wd = ops.pack_conv_weights(w)
    dy = (torch.randn((N, H, W, co), generator=g) * 0.1).to(torch.bfloat16).to(dev)
    dx = torch.empty_like(x)
    cs = torch.empty(ci, dtype=torch.float32, device=dev)
    t0 = timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, 3, 3))
    t1 = timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, 3, 3, relu_mask=x))
    t2 = timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, 3, 3, relu_mask=x, colsum=cs))
    print((N, H, W, ci, co), "dgrad no mask %.1f us, mask %.1f us, mask + colsum %.1f us" % (t0, t1, t2), flush=True)
