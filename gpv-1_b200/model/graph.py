"""CUDA-graph capture of the static-shape training step.

A B=32 step is ~1000 kernel launches of 5-50 us each; issued one by one from Python the host needs ~25 ms to enqueue
them, which is as long as the GPU needs to run them.  `CapturedStep` records the forward (+ criterion) and the
backward of `Engine` once into two CUDA graphs that share a memory pool and replays them per step; per-step data
(images, token ids, ragged targets) is copied into fixed device buffers first, and every batch-dependent scalar
(normalisers of the set criterion) is derived on the device, so a replay is exact for any batch of the captured shape.
The reference has no equivalent (its matcher forces a device->host sync in the middle of every forward: matcher.py:73).
"""
import torch

from .gpv import HostTargets
from .spec import N_STAGES


class CapturedStep:
    def __init__(self, model, images, qids, ans, targets, boxes_per_image_cap=None, warmup=2):
        eng = model.engine
        self.model, self.eng = model, eng
        dev = eng.dev
        B, S = ans.shape
        self.B, self.S, self.Q = B, S, eng.Q
        self.img_shape, self.Tl = tuple(images.shape), qids.shape[1]
        self.train_mode = eng.train_mode                   # the dropout kernels are part of the captured graphs
        cap = boxes_per_image_cap or eng.Q
        self.static_t = HostTargets.alloc_static(B, S, cap, dev)
        self.images = torch.empty(self.img_shape, dtype=images.dtype, device=dev)      # fp32 NCHW or uint8 NHWC
        self.qids = torch.empty((B, self.Tl), dtype=torch.int64, device=dev)
        self.ans = torch.empty((B, S), dtype=torch.int64, device=dev)
        tgt = self._load(images, qids, ans, targets)
        eng.refresh()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        # the warm-up steps are rank-local: no gradient all-reduce.  (A capture must not contain collectives that other ranks may not
        # issue: in a multitask stream each rank captures one step per answer length IT sees, so the number of captures differs
        # between ranks and warm-up all-reduces would pair up wrongly and dead-lock.)
        warm_hooks = (eng.on_stage_done, eng.on_backward_end)
        eng.on_stage_done = eng.on_backward_end = None
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):                   # first-call work (smem attributes, tensor maps, position table)
                    eng.forward_train(self.images, self.qids, self.ans, tgt)
                    eng.backward()
        finally:
            eng.on_stage_done, eng.on_backward_end = warm_hooks
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from .. import _C
        n0 = _C.lib().launches
        self.pool = torch.cuda.graph_pool_handle()
        self.g_fwd = torch.cuda.CUDAGraph()
        # GPVB200_MAIN_PRIO=1: the issuing (critical-path) stream of the captured step gets a higher priority than the lanes, so
        # that its CTAs are placed first whenever both have CTAs waiting for an SM (scheduling experiment; kernel nodes keep the
        # priority of the stream they were captured on)
        import os
        prio = -1 if os.environ.get("GPVB200_MAIN_PRIO", "0") == "1" else 0
        fwd_stream = torch.cuda.Stream(device=dev, priority=prio)
        fwd_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(self.g_fwd, pool=self.pool, stream=fwd_stream):
            self.loss, _ = eng.forward_train(self.images, self.qids, self.ans, tgt)
        self._saved = eng.saved
        # Backward: one graph per gradient stage when a data-parallel GradSync is attached, so that each bucket's NCCL
        # all-reduce is launched (outside any graph, on its side stream) as soon as the stage's graph has been enqueued.
        self.sync = model.grad_sync if (model.grad_sync is not None and model.grad_sync.world > 1) else None
        hooks = (eng.on_stage_done, eng.on_backward_end)
        self.g_bwd = []
        self.g_stage = []                   # gradient stage at which each backward graph ends (where its bucket is reduced)
        cap_stream = torch.cuda.Stream(device=dev, priority=prio)
        cap_stream.wait_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        try:
            with torch.cuda.stream(cap_stream):
                g = torch.cuda.CUDAGraph()
                g.capture_begin(pool=self.pool)
                self.g_bwd.append(g)

                open_capture = [True]

                def split(stage):
                    if self.sync is None or not self.sync.reduces_at(stage):
                        return
                    self.g_stage.append(stage)
                    self.g_bwd[-1].capture_end()
                    if stage == eng.last_stage:           # nothing is launched after the last stage backward reaches: no empty trailing graph
                        open_capture[0] = False
                        return
                    g2 = torch.cuda.CUDAGraph()
                    g2.capture_begin(pool=self.pool)
                    self.g_bwd.append(g2)

                eng.on_stage_done, eng.on_backward_end = split, None
                eng.backward()
                if open_capture[0]:
                    self.g_bwd[-1].capture_end()
        finally:
            eng.on_stage_done, eng.on_backward_end = hooks
        torch.cuda.current_stream().wait_stream(cap_stream)
        torch.cuda.synchronize()
        self.launches_per_step = _C.lib().launches - n0
        self.pending = False

    def matches(self, images, qids, ans):
        return (self.eng.train_mode == self.train_mode and tuple(images.shape) == self.img_shape
                and images.dtype == self.images.dtype and qids.shape[1] == self.Tl
                and tuple(ans.shape) == (self.B, self.S))

    def _load(self, images, qids, ans, targets):
        self.images.copy_(images, non_blocking=True)
        self.qids.copy_(qids, non_blocking=True)
        self.ans.copy_(ans, non_blocking=True)
        return HostTargets(targets, self.B, self.S, self.Q, self.eng.loss_wts, self.eng.eos_coef, self.eng.dev, static=self.static_t)

    def forward(self, images, qids, ans, targets):
        """Copies the batch into the captured buffers and replays the forward graph; returns the loss ([1], static)."""
        tgt = self._load(images, qids, ans, targets)
        if tgt.n_text == 0 and not any("boxes" in t for t in targets):
            return None
        self.eng.refresh()
        self.g_fwd.replay()
        self.pending = True
        return self.loss

    def backward(self):
        assert self.pending, "backward() without a forward()"
        self.pending = False
        if self.sync is None:
            self.g_bwd[0].replay()
            return
        # graph i ends where gradient stage g_stage[i] is complete: its bucket is reduced while the next graph runs
        for g, st in zip(self.g_bwd, self.g_stage):
            g.replay()
            self.sync.stage_done(st)
        self.sync.finish()
