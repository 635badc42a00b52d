"""One FCN-8s training step (B=32, 160x576) for an ncu capture of every tensor-core launch in it.
Plain run: writes the ordered list of tensor-core calls of step 3 to gpurun_out/step_calls.json and
their count K to gpurun_out/step_k.txt.  Under ncu use  -k regex:'^(igemm_kernel|igemm_pair_kernel|wgrad_kernel|wgrad_pair_kernel|slab_kernel|slab3_kernel|wslab_kernel|first_fwd_kernel|first_wgrad_kernel)$'
-s $((2*K)) -c K  (the two warm-up steps launch 2K matching kernels).  Use --metrics + --csv (a
--set full report of ~60 launches exceeds gpurun's 64 MiB return limit)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer
from semanticsegmentation_tensorflow_b200.ops import Profile

TC = ("segk_conv2d_fwd", "segk_conv2d_fwd_pool", "segk_conv2d_dgrad", "segk_conv2d_wgrad", "segk_deconv2d_fwd",
      "segk_deconv2d_dgrad", "segk_deconv2d_wgrad", "segk_deconv2d_packed_fwd", "segk_deconv2d_packed_dgrad",
      "segk_deconv2d_packed_wgrad", "segk_conv2d_first_fwd", "segk_conv2d_first_wgrad")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (32, 160, 576, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 2, (32, 160, 576), dtype=torch.uint8, generator=g).to(dev)
net = FCN(x, 0.8, 2, init="device", overlap=False)
step = AdamOptimizer(1e-4).minimize(net)
feed = {net.image: x, net.annotation: y, net.keep_probability: 0.8}
for _ in range(2):
    step(feed)
torch.cuda.synchronize()
net.ops.profile = Profile()
step(feed)
torch.cuda.synchronize()
calls = [r[0] + str(list(r[5])) for r in net.ops.profile.records if r[0] in TC]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(calls, open("gpurun_out/step_calls.json", "w"), indent=0)
open("gpurun_out/step_k.txt", "w").write(str(len(calls)))
print("tensor-core calls per step:", len(calls))
