"""Batch-sharded data parallelism: one process per GPU, weights replicated, gradients summed
with a bucketed all-reduce that overlaps the rest of backward.

The reference is single-process / single-GPU (`SegNet.py:92-95`); this is new functionality.
The loss gradient is scaled 1/(world * N*H*W) in the xent kernel, so the collective is a
plain SUM over ranks.  Buckets are contiguous slices of the flat gradient arena in backward
completion order; a bucket's all-reduce is issued (async, on the process group's own stream)
as soon as its last layer's weight gradient has been enqueued, and the optimizer update for
that slice waits only on that bucket."""
from __future__ import annotations

import ctypes
import os
from typing import List, Tuple

import torch
import torch.distributed as dist

from . import plan as P


class BucketedAllReduce:
    def __init__(self, flat_grad: torch.Tensor, buckets: List[Tuple[int, int, str]], group=None):
        """flat_grad: the gradient arena (any device); buckets: (start, end, last_layer) in
        backward completion order, e.g. from plan.gradient_buckets()."""
        self.g = flat_grad
        self.buckets = list(buckets)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._next = 0
        self._works = []

    @classmethod
    def for_net(cls, net, group=None, layer_groups=None):
        if hasattr(net, "nodes"):      # GraphNet: generic even-size buckets over its conv/deconv nodes
            names = [n.name for n in net.nodes if n.kind in ("conv", "deconv")]
            return cls(net.vars.g, P.gradient_buckets_even(net.vars.slots, names), group)
        return cls(net.vars.g, P.gradient_buckets(net.vars.slots, layer_groups), group)

    def begin_step(self):
        self._next = 0
        self._works = []

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        if self.world > 1:
            w = dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            w = None
        self._works.append((lo, hi, w))

    def fires_at(self, name: str) -> bool:
        """Does reporting `name` launch a bucket?  (Lets the caller order the launch after the side
        streams only when a collective is actually issued.)"""
        return self._next < len(self.buckets) and self.buckets[self._next][2] == name

    def layer_done(self, name: str):
        """Called by FCN.backward after each layer's dW/db kernels are enqueued.  Returns the buckets
        whose all-reduce was launched by this call as (lo, hi, work) tuples."""
        first = len(self._works)
        while self._next < len(self.buckets) and self.buckets[self._next][2] == name:
            self._launch(self._next)
            self._next += 1
        return self._works[first:]

    def flush(self):
        """Launch the buckets whose completing layer was never reported (defensive); returns them."""
        first = len(self._works)
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        return self._works[first:]

    def finish(self):
        """Yields (lo, hi) arena slices as their reductions complete (stream-ordered wait)."""
        self.flush()
        for lo, hi, w in self._works:
            if w is not None:
                w.wait()
            yield lo, hi


class SymmetricAllReduce(BucketedAllReduce):
    """Same bucket protocol, but the exchange is `segk_allreduce_f32`, our own kernel over NVLink, on a
    gradient arena that lives in symmetric memory: NVLS `multimem.ld_reduce` / `multimem.st` through the
    NVSwitch multicast address when the fabric offers one, peer loads / stores otherwise.  Its 128-thread
    CTAs share SMs with the tensor-core CTAs; an NCCL CTA cannot, which costs the persistent conv kernels
    a second wave while a bucket is in flight (profiles/r1d_scaling.md).  torch's symmetric-memory handle
    provides the allocation, the rendezvous and the cross-rank stream barriers (plumbing); launches are
    stream-ordered on the caller's current stream, so `work` is None."""

    def __init__(self, net, handle, buckets, group=None, param_handle=None):
        super().__init__(net.vars.g, buckets, group)
        self.net = net
        self.ops = net.ops
        self.hdl = handle
        self.rank = dist.get_rank(group)
        self.mc = int(getattr(handle, "multicast_ptr", 0) or 0)        # 0: no multicast on this fabric
        # measured on idle B200s (tools/time_exchange.py, 537 MB): 2 GPUs peer 0.81 ms, multimem 1.34, NCCL 0.99
        want = os.environ.get("SEGK_EXCHANGE", "").lower()      # "", fused, multimem, peer (nccl is handled by try_create)
        if want == "peer":
            self.mc = 0
        self.peers = (ctypes.c_uint64 * self.world)(*[int(p) for p in handle.buffer_ptrs])
        self.kind = "nvls-multimem" if self.mc else "nvlink-peer"
        # fused exchange + Adam (segk_allreduce_adam_f32): needs the parameter arena in symmetric memory too
        self.p_mc = int(getattr(param_handle, "multicast_ptr", 0) or 0) if param_handle is not None else 0
        self.fused = bool(self.mc and self.p_mc) and want != "multimem"
        self._adam = None
        self.slots_stale = False      # fused path: m / v of foreign shares are stale until gather_optimizer_state()
        if self.fused:
            self.kind = "nvls-multimem fused with Adam (sharded optimizer state)"

    def set_adam(self, m, v, lr_t, beta1, beta2, eps):
        """Arms the fused path for this step (TrainStep calls it before backward)."""
        self._adam = (m, v, float(lr_t), float(beta1), float(beta2), float(eps))

    def share(self, lo, hi):
        """[a, b): the slice of arena range [lo, hi) this rank owns in the fused path."""
        n4 = (hi - lo) // 4
        per = -(-n4 // self.world)
        a = lo + 4 * per * self.rank
        return a, max(a, min(a + 4 * per, hi))

    def gather_optimizer_state(self):
        """Fused path: every rank only maintains Adam's m / v for its own shares.  Rebuilds the full slots on
        every rank (e.g. before saving a checkpoint): foreign shares are zeroed, then a SUM all-reduce."""
        V = self.net.vars
        if not self.fused or V.m is None:
            return
        for t in (V.m, V.v):
            keep = torch.zeros_like(t)
            for lo, hi, _ in self.buckets:
                a, b = self.share(lo, hi)
                keep[a:b] = t[a:b]
            dist.all_reduce(keep, op=dist.ReduceOp.SUM, group=self.group)
            t.copy_(keep)
        self.slots_stale = False

    @classmethod
    def try_create(cls, net, group=None):
        """Moves the gradient arena of `net` into symmetric memory and returns the exchange, or None (with
        the reason on stderr) when symmetric memory is unavailable -- the caller then uses NCCL.  Every
        rank takes the same branch: local failures are caught and a MIN all-reduce of the success flag
        decides for all ranks before (and again after) the collective rendezvous."""
        import sys
        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            return None
        mode = os.environ.get("SEGK_EXCHANGE", "").lower()
        if mode == "nccl":
            return None
        dev = net.vars.g.device

        def all_ok(flag: bool) -> bool:
            t = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        def report(err):
            if err is not None:
                print(f"segk: symmetric-memory exchange unavailable ({type(err).__name__}: {err}); using NCCL", file=sys.stderr)

        symm = g = p_new = None
        err = None
        want_params = mode in ("", "fused")     # default: measured best at 2 and 8 GPUs (profiles/r1d_scaling.md)
        try:
            import torch.distributed._symmetric_memory as symm
            g = symm.empty(net.vars.g.numel(), dtype=torch.float32, device=dev)
            g.zero_()
            if want_params:
                # the parameter arena too, so that the exchange can be fused with the optimizer update
                p_new = symm.empty(net.vars.p.numel(), dtype=torch.float32, device=dev)
                p_new.copy_(net.vars.p)
        except Exception as e:       # noqa: BLE001 -- any failure here means "no symmetric memory on this rank"
            err = e
        if not all_ok(err is None):              # some rank could not allocate: nobody enters the rendezvous
            report(err)
            return None
        hdl = phdl = None
        try:
            grp = group if group is not None else dist.group.WORLD
            hdl = symm.rendezvous(g, grp)
            if p_new is not None:
                phdl = symm.rendezvous(p_new, grp)
        except Exception as e:       # noqa: BLE001
            err = e
        if not all_ok(err is None):
            report(err)
            return None
        net.vars.g = g
        if phdl is not None:
            net.vars.p = p_new
        if hasattr(net, "nodes"):
            names = [n.name for n in net.nodes if n.kind in ("conv", "deconv")]
            buckets = P.gradient_buckets_even(net.vars.slots, names)
        else:
            buckets = P.gradient_buckets(net.vars.slots)
        ex = cls(net, hdl, buckets, group, phdl)
        net.exchange = ex            # checkpoint.state_dict gathers the sharded optimizer slots through this
        return ex

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        stream = torch.cuda.current_stream().cuda_stream
        # all ranks' gradients of this bucket are complete (the caller's stream waited for its producers),
        # and no rank still reads the bucket's parameters in this step
        self.hdl.barrier(channel=0, timeout_ms=60000)
        if self.fused and self._adam is not None:
            m, v, lr_t, b1, b2, eps = self._adam
            V = self.net.vars
            self.ops.call("segk_allreduce_adam_f32", self.mc, self.p_mc, V.p.data_ptr(), m.data_ptr(), v.data_ptr(), lo,
                          hi - lo, self.rank, self.world, lr_t, b1, b2, eps, stream)
            self.slots_stale = True
        else:
            self.ops.call("segk_allreduce_f32", self.mc, ctypes.addressof(self.peers), lo, hi - lo, self.rank, self.world,
                          stream)
        self.hdl.barrier(channel=0, timeout_ms=60000)                     # every share has been written everywhere
        self._works.append((lo, hi, None))


def init_distributed(backend: str = "nccl"):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env)."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local
