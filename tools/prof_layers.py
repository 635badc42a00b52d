"""Launch each tensor-core kernel once on the FCN-8s layer shapes (B=32, 160x576), for ncu.
   python tools/prof_layers.py [--time]   (--time: CUDA-event timing of 20 launches each instead)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops, conv_flops

LAYERS = [  # name, N, H, W, Cin, Cout, k
    ("conv1_2", 32, 160, 576, 64, 64, 3),
    ("conv2_2", 32, 80, 288, 128, 128, 3),
    ("conv3_2", 32, 40, 144, 256, 256, 3),
    ("conv4_2", 32, 20, 72, 512, 512, 3),
    ("conv5_2", 32, 10, 36, 512, 512, 3),
    ("conv6", 32, 5, 18, 512, 4096, 7),
    ("conv7", 32, 5, 18, 4096, 4096, 1),
]


def main():
    timing = "--time" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    dev = torch.device("cuda:0")
    ops = Ops(dev)
    out = []
    for name, n, h, w, ci, co, k in LAYERS:
        if only and name not in only:
            continue
        x = torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16)
        dy = torch.randn((n, h, w, co), device=dev).to(torch.bfloat16)
        wt = torch.randn((k, k, ci, co), device=dev) * 0.01
        b = torch.zeros(co, device=dev)
        wk, wd = ops.pack_conv_weights(wt)
        y = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=dev)
        dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=dev)
        dw = torch.empty((k, k, ci, co), dtype=torch.float32, device=dev)
        fl = conv_flops(n, h, w, ci, co, k, k)
        calls = {
            "fwd": lambda: ops.conv2d_fwd(x, wk, b, y, k, k, relu=True),
            "dgrad": lambda: ops.conv2d_dgrad(dy, wd, dx, k, k, relu_mask=x),
            "wgrad": lambda: ops.conv2d_wgrad(x, dy, dw, k, k),
        }
        for op, fn in calls.items():
            if not timing:
                fn()
                continue
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            out.append({"layer": name, "op": op, "ms": ms, "tflops_valid": fl / ms / 1e9,
                        "tflops_dense": 2.0 * n * h * w * k * k * ci * co / ms / 1e9})
            print(f"{name:8s} {op:6s} {ms:8.4f} ms  {fl / ms / 1e9:8.1f} TF/s valid  "
                  f"{2.0 * n * h * w * k * k * ci * co / ms / 1e9:8.1f} TF/s dense", flush=True)
        del x, dy, wt, wk, wd, y, dx, dw
    torch.cuda.synchronize()
    if timing:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/prof_layers.json", "w"), indent=1)


if __name__ == "__main__":
    main()
