"""Diagnostic for the tcgen05 path: structured inputs whose expected output is obvious, with
a dump of what came back.  Run on the GPU box when a parity test fails."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda:0")
ops = Ops(dev)
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)


def run_1x1(n, h, w, ci, co):
    # x[p, c] = p % 7 + c/64 ;  W = identity-ish: W[ci, co] = 1 if ci == co % ci_total
    npix = n * h * w
    x = (torch.arange(npix).view(-1, 1) % 7).float() + torch.arange(ci).view(1, -1).float() / 64
    wt = torch.zeros(1, 1, ci, co)
    for o in range(co):
        wt[0, 0, o % ci, o] = 1.0
    xb = x.to(torch.bfloat16)
    ref = (xb.float() @ wt[0, 0]).view(n, h, w, co)
    wk, wd = ops.pack_conv_weights(wt.to(dev))
    y = torch.full((n, h, w, co), -77.0, dtype=torch.bfloat16, device=dev)
    ops.conv2d_fwd(xb.view(n, h, w, ci).to(dev), wk, None, y, 1, 1, relu=False)
    torch.cuda.synchronize()
    got = y.float().cpu()
    err = (got - ref).abs()
    print(f"1x1 n{n} h{h} w{w} ci{ci} co{co}: max err {err.max():.4f}, mismatches {(err > 0.05).sum().item()} / {err.numel()}")
    if err.max() > 0.05:
        bad = (err > 0.05).nonzero()
        print(" first bad idx", bad[:8].tolist())
        print(" got[0,0,0,:16]", got[0, 0, 0, :16])
        print(" ref[0,0,0,:16]", ref[0, 0, 0, :16])
        print(" got[0,0,1,:16]", got[0, 0, 1, :16])
        print(" ref[0,0,1,:16]", ref[0, 0, 1, :16])
        rows_bad = (err.view(npix, co) > 0.05).any(1).nonzero().flatten()
        cols_bad = (err.view(npix, co) > 0.05).any(0).nonzero().flatten()
        print(" bad rows", rows_bad[:40].tolist(), "n", rows_bad.numel())
        print(" bad cols", cols_bad[:40].tolist(), "n", cols_bad.numel())


def run_wgrad(n, h, w, ci, co):
    npix = n * h * w
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.integers(-2, 3, (npix, ci)).astype(np.float32))
    dy = torch.tensor(rng.integers(-2, 3, (npix, co)).astype(np.float32))
    ref = x.t() @ dy
    dw = torch.zeros((1, 1, ci, co), dtype=torch.float32, device=dev)
    ops.conv2d_wgrad(x.view(n, h, w, ci).to(torch.bfloat16).to(dev), dy.view(n, h, w, co).to(torch.bfloat16).to(dev), dw, 1, 1)
    torch.cuda.synchronize()
    got = dw.cpu()[0, 0]
    err = (got - ref).abs()
    print(f"wgrad 1x1 n{n} h{h} w{w} ci{ci} co{co}: max err {err.max():.4f} mismatches {(err > 0.01).sum().item()} / {err.numel()}")
    if err.max() > 0.01:
        print(" got[:4,:8]", got[:4, :8])
        print(" ref[:4,:8]", ref[:4, :8])
        rows_bad = (err > 0.01).any(1).nonzero().flatten()
        cols_bad = (err > 0.01).any(0).nonzero().flatten()
        print(" bad rows", rows_bad[:40].tolist(), "n", rows_bad.numel())
        print(" bad cols", cols_bad[:40].tolist(), "n", cols_bad.numel())


if __name__ == "__main__":
    run_1x1(1, 8, 16, 64, 64)
    run_1x1(1, 8, 16, 128, 128)
    run_1x1(2, 16, 16, 64, 256)
    run_1x1(1, 4, 8, 64, 64)      # partial box (32 rows)
    run_wgrad(1, 8, 8, 64, 64)
    run_wgrad(1, 8, 16, 128, 128)
    run_wgrad(2, 16, 16, 64, 256)
    run_wgrad(2, 16, 16, 192, 64)
    print("diag done")


def run_slab(bo_mode, n=1, h=8, w=64, ci=64, co=64):
    """3x3 conv through the slab kernel with integer data; prints the error under a base-offset mode."""
    ops.ctx.set_tuning("slab", 2)
    rng = np.random.default_rng(1)
    x = torch.tensor(rng.integers(-2, 3, (n, h, w, ci)).astype(np.float32))
    wt = torch.tensor(rng.integers(-1, 2, (3, 3, ci, co)).astype(np.float32))
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), wt.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    wk, _ = ops.pack_conv_weights(wt.to(dev))
    y = torch.full((n, h, w, co), -77.0, dtype=torch.float32, device=dev)
    ops.conv2d_fwd(x.to(torch.bfloat16).to(dev), wk, None, y, 3, 3, relu=False)
    torch.cuda.synchronize()
    err = (y.cpu() - ref).abs()
    print(f"slab bo_mode={bo_mode} n{n} h{h} w{w} ci{ci} co{co}: max err {err.max():.3f} mismatches {(err > 0.01).sum().item()} / {err.numel()}")
    if err.max() > 0.01:
        bad = (err > 0.01).any(-1)
        print("  bad pixel map (n=0), rows = y:")
        for yy in range(min(h, 8)):
            print("   ", "".join("X" if bad[0, yy, xx] else "." for xx in range(w)))
    ops.ctx.set_tuning("slab", 1)


if __name__ == "__main__":
    run_slab(0)
    run_slab(0, 2, 16, 96, 128, 128)
