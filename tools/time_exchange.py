"""torchrun worker: all-reduce of the 537 MB fp32 gradient arena, NCCL vs segk_allreduce_f32 (idle GPUs)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from semanticsegmentation_tensorflow_b200.dp import init_distributed
from semanticsegmentation_tensorflow_b200.ops import Ops

rank, world, local = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
ops = Ops(dev)
n = 134_217_728
g = symm.empty(n, dtype=torch.float32, device=dev)
g.fill_(1.0)
hdl = symm.rendezvous(g, dist.group.WORLD)
mc = int(hdl.multicast_ptr or 0)
peers = (ctypes.c_uint64 * world)(*[int(p) for p in hdl.buffer_ptrs])
x = torch.ones(n, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ours(use_mc):
    hdl.barrier(channel=0)
    ops.call("segk_allreduce_f32", mc if use_mc else 0, ctypes.addressof(peers), 0, n, rank, world,
             torch.cuda.current_stream().cuda_stream)
    hdl.barrier(channel=0)


t_nccl = timeit(lambda: dist.all_reduce(x))
res = {"nccl_ms": t_nccl}
if mc:
    res["multimem_ms"] = timeit(lambda: ours(True))
res["peer_ms"] = timeit(lambda: ours(False))
res["barrier_pair_ms"] = timeit(lambda: (hdl.barrier(channel=0), hdl.barrier(channel=0)))
if rank == 0:
    gb = n * 4 / 1e9
    print({k: round(v, 3) for k, v in res.items()}, "algbw GB/s:", {k: round(gb / (v / 1e3), 1) for k, v in res.items() if "barrier" not in k})
dist.barrier()
dist.destroy_process_group()
