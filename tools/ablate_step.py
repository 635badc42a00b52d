"""Step-time ablation (B=32, 160x576 FCN-8s): what the side-stream work costs the main stream.
Variants: full step; without BiasAddGrad; without Adam + repack; without both; forward only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (32, 160, 576, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 2, (32, 160, 576), dtype=torch.uint8, generator=g).to(dev)
net = FCN(x, 0.8, 2, init="device")
opt = AdamOptimizer(1e-4)
step = opt.minimize(net)
feed = {net.image: x, net.annotation: y, net.keep_probability: 0.8}


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


ops = net.ops
print("full step            %.3f ms" % timeit(lambda: step(feed)))
real_bias = ops.bias_grad
ops.bias_grad = lambda d, g: g
print("no bias_grad         %.3f ms" % timeit(lambda: step(feed)))
real_apply, real_repack, real_fused = opt.apply, net.vars.repack, opt.apply_and_repack
opt.apply = lambda *a, **k: None
opt.apply_and_repack = lambda *a, **k: None
net.vars.repack = lambda *a, **k: None
print("no bias, adam, pack  %.3f ms" % timeit(lambda: step(feed)))
ops.bias_grad = real_bias
print("no adam, pack        %.3f ms" % timeit(lambda: step(feed)))
opt.apply, net.vars.repack, opt.apply_and_repack = real_apply, real_repack, real_fused


def fwd_only():
    net.feed(feed)
    net.forward()
    net.loss(with_grad=True)


print("forward + loss       %.3f ms" % timeit(fwd_only))


def fwd_bwd():
    fwd_only()
    net.backward()
    net.side.join()
    net.wside.join()


print("fwd + bwd (no opt)   %.3f ms" % timeit(fwd_bwd))
net.side.enabled = net.wside.enabled = False
print("fwd + bwd serial     %.3f ms" % timeit(fwd_bwd))
ops.bias_grad = lambda d, g: g
print("fwd + bwd serial, no bias_grad %.3f ms" % timeit(fwd_bwd))
