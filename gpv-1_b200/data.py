"""Input staging for the training loop: double-buffered pinned-host -> device copies on their own stream.

The reference moves each batch to the GPU synchronously inside the loop (`imgs.to(torch.device(gpu))`,
exp/gpv/train_distr.py:401), so 118 MB of fp32 pixels cross PCIe (~2.5 ms) before the step can start.  `DevicePrefetcher`
wraps any iterable of `(images, queries, targets)` batches: while step i computes, batch i+1 is copied on a side stream
into the second of two device buffers; `next()` makes the compute stream wait for that copy's event and yields device
tensors, which `GPV.forward` consumes without another copy.  Images may be fp32 NCHW or the loader's raw uint8 NHWC.
"""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device):
        self.it = iter(batches)
        self.dev = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.buf = [None, None]        # device image buffers (ping-pong), allocated on first use per shape/dtype
        self.slot = 0
        self.ready = None
        self._stage()

    def _pin(self, t):
        return t if (not torch.is_tensor(t) or t.is_cuda or t.is_pinned()) else t.pin_memory()

    def _stage(self):
        try:
            images, queries, targets = next(self.it)
        except StopIteration:
            self.ready = None
            return
        s = self.slot
        self.slot ^= 1
        if torch.is_tensor(images) and not images.is_cuda:
            if self.buf[s] is None or self.buf[s].shape != images.shape or self.buf[s].dtype != images.dtype:
                self.buf[s] = torch.empty(images.shape, dtype=images.dtype, device=self.dev)
            src = self._pin(images)
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))   # the buffer's previous consumer has been enqueued
            with torch.cuda.stream(self.stream):
                self.buf[s].copy_(src, non_blocking=True)
                if torch.is_tensor(queries):
                    queries = self._pin(queries).to(self.dev, non_blocking=True)
                targets = [{k: (self._pin(v).to(self.dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in t.items()}
                           for t in targets]
            images = self.buf[s]
            keep = src
        else:
            keep = None
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.ready = (images, queries, targets, ev, keep)

    def __iter__(self):
        return self

    def __next__(self):
        if self.ready is None:
            raise StopIteration
        images, queries, targets, ev, _keep = self.ready
        torch.cuda.current_stream(self.dev).wait_event(ev)
        self._stage()                  # batch i+1 starts crossing PCIe while batch i computes
        return images, queries, targets
