"""Launch the conv3_2 / conv4_2 / conv6 forward convs as single-CTA tiles and as CTA pairs (for an ncu capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda", 0)
ops = Ops(dev)
for n, h, w, ci, co, k in [(32, 40, 144, 256, 256, 3), (32, 20, 72, 512, 512, 3), (32, 5, 18, 512, 4096, 7)]:
    x = torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16)
    wt = torch.randn((k, k, ci, co), device=dev) * 0.02
    b = torch.zeros(co, device=dev)
    wk, wd = ops.pack_conv_weights(wt)
    y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=dev)
    for mode in (0, 1):
        ops.ctx.set_tuning("pair", mode)
        for _ in range(2):
            ops.conv2d_fwd(x, wk, b, y, k, k, relu=True)
        torch.cuda.synchronize()
ops.ctx.set_tuning("pair", 1)
