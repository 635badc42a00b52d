"""A/B of the transposed conv forward's epilogue: direct strided stores vs TMA store of per-phase decimated views."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops
dev = torch.device("cuda", 0)
ops = Ops(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for name, n, h, w, ci, co in [("unpool1 5x18 512->512", 32, 5, 18, 512, 512), ("unpool2 10x36 512->512", 32, 10, 36, 512, 512),
                              ("unpool3 20x72 256->256", 32, 20, 72, 256, 256), ("unpool4 40x144 128->128", 32, 40, 144, 128, 128),
                              ("unpool5 80x288 64->64", 32, 80, 288, 64, 64)]:
    x = torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16)
    wk, _ = ops.pack_deconv_weights(torch.randn((4, 4, co, ci), device=dev) * 0.02, 2)
    y = torch.empty((n, 2 * h, 2 * w, co), dtype=torch.bfloat16, device=dev)
    r = {}
    for mode in (0, 1, 0, 1):
        ops.ctx.set_tuning("tma_store", mode)
        r.setdefault(mode, []).append(timeit(lambda: ops.deconv2d_fwd(x, wk, None, y, 4, 2)))
    ops.ctx.set_tuning("tma_store", 1)
    print(f"{name:26s} direct {min(r[0]):.1f} us | tma store {min(r[1]):.1f} us", flush=True)
