"""Data-parallel gradient exchange for the GPV step (replaces DistributedDataParallel at train_distr.py:192-193).

The reference wraps the model in DDP(find_unused_parameters=True): 227 M registered parameters, 25 MB fp32 buckets
discovered by autograd hooks, an unused-parameter graph walk every step.  Here the engine already owns ONE flat fp32
gradient arena laid out in the order backward completes it (model/spec.py:grad_stage -- text decoder, co-attention,
DETR decoder, DETR encoder, layer4, layer3, layer2) and holding only the ~114 M parameters that can receive a
gradient.  `GradSync` turns stage boundaries into buckets (by default after the DETR encoder, layer4, layer3 and layer2: four
contiguous prefixes of the arena): when backward finishes such a stage it records an event and the bucket's all-reduce (NCCL over
NVLink/NVSwitch) is enqueued on a side stream, so every bucket but the last (layer2, ~1.2 M elements) is reduced underneath the
remaining backward kernels.  There is no other collective on the
path: the criterion is rank-local (set_criterion.py:165-168 has the num_boxes all-reduce commented out) and clipping
happens after averaging on identical replicas (train_distr.py:423-426).
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, model=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.engine = None
        self.stream = None
        self.bytes_per_step = 0
        # Buckets are reduced after gradient stages 3, 4, 5, 6 (DETR encoder, layer4, layer3, layer2) and after the last stage backward
        # reaches; each bucket spans every stage completed since the previous one.  The first bucket (text decoder + co-attention +
        # DETR decoder + encoder, 364 MB) then starts half-way through backward and hides under the trunk's 4 ms; three fewer graph
        # boundaries in the captured backward (N = 2: 16.44 -> 16.34 ms per step, profiles/r3e_ddp_buckets.txt).
        # GPVB200_DDP_REDUCE_AT=all restores one bucket per stage.
        import os
        ra = os.environ.get("GPVB200_DDP_REDUCE_AT", "3,4,5,6")      # "all" = after every stage
        self.reduce_at = None if ra.strip().lower() == "all" else ({int(x) for x in ra.split(",") if x.strip() != ""} or None)
        self._lo = 0                       # arena offset where the next bucket starts
        if model is not None:
            model.grad_sync = self
            if getattr(model, "_engine", None) is not None:
                self.attach(model._engine)

    def attach(self, engine):
        self.engine = engine
        engine.on_stage_done = self.stage_done
        engine.on_backward_end = self.finish
        engine.stage_joins = self.reduces_at                   # the engine joins its lanes only where a bucket ends
        self.cuda = engine.grad_arena.is_cuda
        self.stream = torch.cuda.Stream(device=engine.grad_arena.device) if self.cuda else None
        self.bytes_per_step = engine.grad_arena.numel() * 4

    def bucket(self, stage):
        e = self.engine.stage_end
        return (e[stage - 1] if stage else 0), e[stage]

    def reduces_at(self, stage):
        return self.reduce_at is None or stage in self.reduce_at or stage == getattr(self.engine, "last_stage", None)

    def stage_done(self, stage):
        if self.world == 1:
            return
        if not self.reduces_at(stage):
            return
        lo, hi = (self._lo if self.reduce_at is not None else self.bucket(stage)[0]), self.bucket(stage)[1]
        self._lo = hi
        if hi <= lo:
            return
        buf = self.engine.grad_arena[lo:hi]
        if not self.cuda:                       # gloo / CPU tests: same bucket schedule, synchronous
            buf.div_(self.world)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)

    def finish(self):
        """Called at the end of backward: the compute stream waits for the outstanding bucket reductions."""
        self._lo = 0
        if self.world > 1 and self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)


def broadcast_parameters(model, src=0, group=None):
    """Rank `src`'s parameters and buffers to every replica (DDP does this at construction)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in model.state_dict().values():
        dist.broadcast(t, src=src, group=group)
