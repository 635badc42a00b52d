"""torchrun worker for tests/test_gpu_dp.py: batch-sharded data parallelism over NCCL.
Each rank trains on its shard of a fixed global batch; rank 0 also runs the same global batch on one
GPU; after 3 steps the replicas must be bit-identical across ranks and close to the single-GPU run."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from semanticsegmentation_tensorflow_b200.dp import BucketedAllReduce, SymmetricAllReduce, init_distributed
from semanticsegmentation_tensorflow_b200.fcn import FCN, AdamOptimizer, reference_init
from semanticsegmentation_tensorflow_b200 import plan as P

rank, world, local = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
FC, H, W, GB = 128, 64, 96, 8
rng = np.random.default_rng(0)
x = (rng.integers(0, 256, (GB, H, W, 3)) // 32).astype(np.uint8)
y = rng.integers(0, 2, (GB, H, W)).astype(np.uint8)
variables = reference_init(P.variable_shapes(3, 2, FC), 1234, "he")
lo, hi = P.shard_batch(GB, world, rank)
xs, ys = torch.as_tensor(x[lo:hi]).to(dev), torch.as_tensor(y[lo:hi]).to(dev)
net = FCN(xs, 1.0, 2, variables=variables, fc=FC, world_size=world)
MODE = os.environ.get("DP_EXCHANGE", "nccl")
if MODE in ("symmetric", "fused"):   # our NVLink kernels on symmetric-memory arenas; fused = exchange + Adam in one kernel
    os.environ["SEGK_EXCHANGE"] = "fused" if MODE == "fused" else "multimem"
    ar = SymmetricAllReduce.try_create(net)
    if ar is None:
        if rank == 0:
            print("DPRESULT " + json.dumps({"skipped": "no symmetric memory on this box"}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0)
else:
    ar = BucketedAllReduce.for_net(net)
# (1) the reduced gradient itself (Adam is scale-invariant, so check it before any update)
net.forward()
net.loss(ys, with_grad=True)
ar.begin_step()
net.backward(after_layer=lambda name: (net.side.join(), net.wside.join(), ar.layer_done(name)))
net.side.join()
net.wside.join()
for _ in ar.finish():
    pass
torch.cuda.synchronize()
g_dp = net.vars.g.clone()
step = AdamOptimizer(1e-4).minimize(net, allreduce=ar)
losses = []
for _ in range(3):
    l = step({net.image: xs, net.annotation: ys, net.keep_probability: 1.0})
    t = l.detach().clone()
    dist.all_reduce(t)                     # mean over ranks of per-shard mean losses = global mean loss
    losses.append(float(t) / world)
torch.cuda.synchronize()
if getattr(ar, "fused", False):
    ar.gather_optimizer_state()           # every rank holds the full Adam slots again
    m0 = net.vars.m.clone()
    dist.broadcast(m0, 0)
    assert float((net.vars.m - m0).abs().max()) == 0.0
# replicas identical: max |p - p_rank0| == 0
p0 = net.vars.p.clone()
dist.broadcast(p0, 0)
same = float((net.vars.p - p0).abs().max())
gather = [None] * world
dist.all_gather_object(gather, same)
out = {"rank": rank, "replica_max_diff": max(gather), "losses": losses, "exchange": getattr(ar, "kind", "nccl")}
if rank == 0:
    xf, yf = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
    ref = FCN(xf, 1.0, 2, variables=variables, fc=FC, world_size=1)
    ref.forward()
    ref.loss(yf, with_grad=True)
    ref.backward()
    torch.cuda.synchronize()
    g1 = ref.vars.g
    out["grad_cosine"] = float(torch.dot(g_dp, g1) / (g_dp.norm() * g1.norm()))
    out["grad_norm_ratio"] = float(g_dp.norm() / g1.norm())
    rstep = AdamOptimizer(1e-4).minimize(ref)
    out["single_losses"] = [float(rstep({ref.image: xf, ref.annotation: yf, ref.keep_probability: 1.0})) for _ in range(3)]
    torch.cuda.synchronize()
    d = (net.vars.p - ref.vars.p).abs()
    out["vs_single_max"], out["vs_single_mean"] = float(d.max()), float(d.mean())
    print("DPRESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
