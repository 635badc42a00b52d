"""Single GPU: step time through the local-bucket path vs the data-parallel code path with the
collective stubbed out (world size 1) -- separates host/stream-structure overhead of the DP path from
the cost of the NCCL kernels themselves."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer, TrainStep
from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (32, 160, 576, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 2, (32, 160, 576), dtype=torch.uint8, generator=g).to(dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


net = FCN(x, 0.8, 2, init="device")
opt = AdamOptimizer(1e-4)
feed = {net.image: x, net.annotation: y, net.keep_probability: 0.8}
local = opt.minimize(net)
print("local-bucket path      %.3f ms" % timeit(lambda: local(feed)))
dp = TrainStep(net, opt, allreduce=BucketedAllReduce.for_net(net))
print("DP path, no collective %.3f ms" % timeit(lambda: dp(feed)))
