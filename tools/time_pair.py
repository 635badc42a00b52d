"""A/B timing of the 256-column igemm launches: CTA pairs (cta_group::2) vs single-CTA tiles, same box, interleaved.
Usage: python tools/time_pair.py  (prints one line per layer shape: us single, us pair)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda", 0)
ops = Ops(dev)
SHAPES = [  # name, N, H, W, Cin, Cout, k
    ("conv2_1 80x288 64->128 (slab)", 32, 80, 288, 64, 128, 3),
    ("conv2_2 80x288 128->128 (slab)", 32, 80, 288, 128, 128, 3),
    ("conv3_1 40x144 128->256", 32, 40, 144, 128, 256, 3),
    ("conv3_2 40x144 256->256", 32, 40, 144, 256, 256, 3),
    ("conv4_1 20x72 256->512", 32, 20, 72, 256, 512, 3),
    ("conv4_2 20x72 512->512", 32, 20, 72, 512, 512, 3),
    ("conv5_x 10x36 512->512", 32, 10, 36, 512, 512, 3),
    ("conv6 5x18 512->4096 7x7", 32, 5, 18, 512, 4096, 7),
    ("conv7 5x18 4096->4096 1x1", 32, 5, 18, 4096, 4096, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, n, h, w, ci, co, k in SHAPES:
    x = torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16)
    dy = torch.randn((n, h, w, co), device=dev).to(torch.bfloat16)
    wt = torch.randn((k, k, ci, co), device=dev) * 0.02
    b = torch.zeros(co, device=dev)
    wk, wd = ops.pack_conv_weights(wt)
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=dev)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=dev)
    xb = torch.empty((n, h, w, ci // 32), dtype=torch.int32, device=dev)
    dw = torch.empty((k, k, ci, co), dtype=torch.float32, device=dev)
    ops.relu_bits(x, xb)
    res = {}
    for rep in range(2):
        for mode in (0, 1):
            ops.ctx.set_tuning("pair", mode)
            f = timeit(lambda: ops.conv2d_fwd(x, wk, b, y, k, k, relu=True))
            d = timeit(lambda: ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask_bits=xb))
            g = timeit(lambda: ops.conv2d_wgrad(x, dy, dw, k, k))
            res.setdefault(mode, []).append((f, d, g))
    ops.ctx.set_tuning("pair", 1)
    s = " | ".join("pair=%d fwd %.1f dgrad %.1f wgrad %.1f" % (m, min(r[0] for r in res[m]), min(r[1] for r in res[m]), min(r[2] for r in res[m]))
                   for m in (0, 1))
    print(f"{name:28s} {s}", flush=True)

# the stride-2 implicit GEMM (input gradient of the 4x4 / stride-2 transposed convs of U-Net / SegNet)
for name, n, h, w, ci, co in [("unpool1 dgrad 5x18 512", 32, 5, 18, 512, 512), ("unpool2 dgrad 10x36 512", 32, 10, 36, 512, 512),
                              ("unpool3 dgrad 20x72 256", 32, 20, 72, 256, 256)]:
    wt = torch.randn((4, 4, co, ci), device=dev) * 0.02
    _, wd = ops.pack_deconv_weights(wt, 2)
    dy = torch.randn((n, 2 * h, 2 * w, co), device=dev).to(torch.bfloat16)
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=dev)
    res = {}
    for rep in range(2):
        for mode in (0, 1):
            ops.ctx.set_tuning("pair", mode)
            res.setdefault(mode, []).append(timeit(lambda: ops.deconv2d_dgrad(dy, wd, dx, 4, 2)))
    ops.ctx.set_tuning("pair", 1)
    print(f"{name:28s} pair=0 {min(res[0]):.1f} | pair=1 {min(res[1]):.1f}", flush=True)
