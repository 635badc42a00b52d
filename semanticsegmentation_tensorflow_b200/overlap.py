"""Side CUDA streams.  (1) For the HBM-bound work that is off the tensor-core critical path of a training
step: BiasAddGrad of each layer, and the optimizer update + bf16 weight repack of a gradient bucket as
soon as that bucket is complete.  The persistent tensor-core kernels hold one 192-thread CTA per SM
(~200 KB of shared memory), so these streaming kernels co-reside on the same SMs and run in the
shadow of the conv GEMMs instead of after them.  (2) A second one carries the weight-gradient GEMMs:
they are off the dgrad -> dgrad critical path of backward, so running them beside the next input-
gradient kernels lets their CTAs fill the SMs a kernel's last partial wave leaves idle (conv5: 180
tiles on 148 SMs = 61 % of two waves)."""
from __future__ import annotations

import torch


class SideStream:
    def __init__(self, device, enabled=True):
        self.enabled = enabled
        self.stream = torch.cuda.Stream(device) if enabled else None
        self._reads = {}          # storage ptr -> event after the last side-stream read of that buffer

    def mark(self):
        """Event on the current stream now; pass it to run(after=...) to order side work after this
        point only (and not after what the main stream enqueues in between)."""
        if not self.enabled:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    def run(self, fn, reads=(), after=None):
        """Enqueue fn() on the side stream, ordered after everything enqueued on the current stream so
        far (or up to the mark() event `after`).  `reads`: tensors the side work reads that the main
        stream may later overwrite."""
        if not self.enabled:
            fn()
            return
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            fn()
        if reads:
            done = torch.cuda.Event()
            done.record(self.stream)
            for t in reads:
                self._reads[t.untyped_storage().data_ptr()] = done

    def before_write(self, t):
        """Call before the main stream overwrites tensor t."""
        if not self.enabled:
            return
        ev = self._reads.pop(t.untyped_storage().data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def join(self):
        if self.enabled:
            torch.cuda.current_stream().wait_stream(self.stream)
            self._reads.clear()


class LocalBuckets:
    """Single-GPU counterpart of dp.BucketedAllReduce: when a gradient bucket is complete, its optimizer
    update and weight repack are issued on the side stream while backward continues."""

    def __init__(self, net, opt, buckets):
        self.net, self.opt, self.buckets = net, opt, list(buckets)
        slots = net.vars.slots
        self.layers = []
        for lo, hi, _ in self.buckets:
            self.layers.append({n.split("/")[0] for n, s in slots.items() if lo <= s.offset < hi})
        self._next = 0

    def begin_step(self):
        self._next = 0

    def _fire(self, b):
        lo, hi, _ = self.buckets[b]
        names = self.layers[b]
        net, opt = self.net, self.opt

        def work():
            if hasattr(opt, "apply_and_repack"):
                opt.apply_and_repack(net, lo, hi, names)     # Adam fused with the bf16 repack where a layer allows it
            else:
                opt.apply(net, lo, hi)
                net.vars.repack(net.ops, names)

        wside = getattr(net, "wside", None)
        if wside is not None and wside.enabled and net.side.enabled:
            net.side.stream.wait_stream(wside.stream)     # the bucket's weight gradients come from there
        net.side.run(work)

    def layer_done(self, name):
        while self._next < len(self.buckets) and self.buckets[self._next][2] == name:
            self._fire(self._next)
            self._next += 1

    def finish(self):
        while self._next < len(self.buckets):
            self._fire(self._next)
            self._next += 1
        self.net.side.join()
        if getattr(self.net, "wside", None) is not None:
            self.net.wside.join()
