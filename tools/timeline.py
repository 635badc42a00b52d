"""Poor man's timeline of one training step (no nsys in the image): CUDA events on the main stream at
every layer of backward and around every item of the side / wgrad streams, printed relative to the start
of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer
from semanticsegmentation_tensorflow_b200 import overlap

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (32, 160, 576, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 2, (32, 160, 576), dtype=torch.uint8, generator=g).to(dev)
net = FCN(x, 0.8, 2, init="device")
opt = AdamOptimizer(1e-4)
step = opt.minimize(net)
feed = {net.image: x, net.annotation: y, net.keep_probability: 0.8}
for _ in range(3):
    step(feed)
torch.cuda.synchronize()

records = []          # (stream name, label, start event, end event)
main_marks = []       # (label, event)


def ev():
    return torch.cuda.Event(enable_timing=True)


def wrap(side, sname):
    orig = side.run

    def run(fn, reads=(), after=None, _label=[0]):
        def timed():
            e0, e1 = ev(), ev()
            e0.record()
            fn()
            e1.record()
            records.append((sname, _label[0], e0, e1))
        _label[0] += 1
        return orig(timed, reads=reads, after=after)

    side.run = run


wrap(net.side, "side")
wrap(net.wside, "wgrad")
orig_fire = step.local._fire


def fire(b):
    e = ev(); e.record(); main_marks.append((f"fire bucket {b}", e))
    orig_fire(b)


step.local._fire = fire
orig_layer_done = step.local.layer_done


def layer_done(name):
    e = ev(); e.record(); main_marks.append((f"bwd {name} enqueued", e))
    orig_layer_done(name)


step.local.layer_done = layer_done
t0 = ev(); t0.record()
net.feed(feed)
net.forward()
loss = net.loss(with_grad=True)
tf = ev(); tf.record()
opt.t += 1
step.local.begin_step()
net.backward(after_layer=step.local.layer_done)
tb = ev(); tb.record()
step.local.finish()
te = ev(); te.record()
torch.cuda.synchronize()
print(f"forward+loss done {t0.elapsed_time(tf):.3f} ms; backward (main stream) done {t0.elapsed_time(tb):.3f}; step done {t0.elapsed_time(te):.3f}")
for label, e in main_marks:
    print(f"  main  {t0.elapsed_time(e):7.3f}  {label}")
for sname, label, e0, e1 in records:
    print(f"  {sname:5s} {t0.elapsed_time(e0):7.3f} -> {t0.elapsed_time(e1):7.3f}  ({e0.elapsed_time(e1):.3f} ms) item {label}")
