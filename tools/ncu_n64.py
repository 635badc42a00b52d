"""The four 64-channel full-resolution launches of a training step (conv1_1 fused fwd / wgrad,
conv1_2 slab fwd, conv1_2 wgrad) at B=32 160x576, three times each, for an ncu --set full capture
with source:  -k regex:'first_fwd_kernel|first_wgrad_kernel|slab_kernel|wgrad_kernel' -s 8 -c 4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.ops import Ops

dev = torch.device("cuda:0")
ops = Ops(dev)
N, H, W = 32, 160, 576
g = torch.Generator().manual_seed(0)
img = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
w1 = (torch.randn((3, 3, 3, 64), generator=g) * 0.01).to(dev)
w2 = (torch.randn((3, 3, 64, 64), generator=g) * 0.01).to(dev)
b = torch.zeros(64, device=dev)
wk1 = ops.pack_im2col_weights(w1)
wk2, wd2 = ops.pack_conv_weights(w2)
a1 = torch.empty((N, H, W, 64), dtype=torch.bfloat16, device=dev)
a2 = torch.empty_like(a1)
dy = (torch.randn((N, H, W, 64), generator=g) * 0.1).to(torch.bfloat16).to(dev)
dw1 = torch.empty_like(w1)
db1 = torch.empty(64, device=dev)
dw2 = torch.empty_like(w2)
for _ in range(3):
    ops.conv2d_first_fwd(img, wk1, b, a1, 3, 3, relu=True)
    ops.conv2d_first_wgrad(img, dy, dw1, 3, 3, dbias=db1)
    ops.conv2d_fwd(a1, wk2, b, a2, 3, 3, relu=True)
    ops.conv2d_wgrad(a1, dy, dw2, 3, 3)
torch.cuda.synchronize()
print("ok")
