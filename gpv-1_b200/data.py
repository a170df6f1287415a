"""Input staging for the training loop: double-buffered pinned-host -> device copies on their own stream.

The reference moves each batch to the GPU synchronously inside the loop (`imgs.to(torch.device(gpu))`,
exp/gpv/train_distr.py:401), so 118 MB of fp32 pixels cross PCIe (~2.5 ms) before the step can start.  `DevicePrefetcher`
wraps any iterable of `(images, queries, targets)` batches: while step i computes, batch i+1 is copied on a side stream
into the second of two device buffers; `next()` makes the compute stream wait for that copy's event and yields device
tensors, which `GPV.forward` consumes without another copy.  Images may be fp32 NCHW or the loader's raw uint8 NHWC.
"""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device, move_targets=False):
        """move_targets=False leaves the per-sample target dicts on the host (pinned): GPV.forward packs them into ONE staging buffer
        and one copy (HostTargets._stage_host), which is cheaper than 2-3 tiny tensors per sample moved here and re-assembled by
        a dozen small device kernels in front of every step (0.6 ms per step at B = 32)."""
        self.move_targets = move_targets
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
        keep = None
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))   # the buffers' previous consumer has been enqueued
        with torch.cuda.stream(self.stream):
            if torch.is_tensor(images) and not images.is_cuda:
                if self.buf[s] is None or self.buf[s].shape != images.shape or self.buf[s].dtype != images.dtype:
                    self.buf[s] = torch.empty(images.shape, dtype=images.dtype, device=self.dev)
                keep = self._pin(images)
                self.buf[s].copy_(keep, non_blocking=True)
                images = self.buf[s]
            # (a list of differently sized images is passed through: GPV.forward pads and copies it itself)
            if torch.is_tensor(queries):
                queries = self._pin(queries).to(self.dev, non_blocking=True)
            if self.move_targets:
                targets = [{k: (self._pin(v).to(self.dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in t.items()}
                           for t in targets]
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


class LossReader:
    """Per-step scalar read-back that never drains the GPU queue.  `loss.item()` right after `backward()` (the reference's logging,
    train_distr.py:433) makes the host wait for the step and only then enqueue the next one: ~1 ms of idle GPU per 14 ms step
    (host enqueue of the captured graphs + the synchronisation round trip).  `push(loss)` enqueues a 4-byte device -> pinned-host
    copy behind the step and returns the PREVIOUS step's value, whose copy finished long ago; `flush()` returns the last one.
    Every step's loss still reaches the host, one step late."""

    def __init__(self):
        self.host = torch.empty(2, dtype=torch.float32).pin_memory()
        self.ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0
        self.pending = False

    def push(self, loss):
        s = self.slot
        self.host[s:s + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.ev[s].record()
        prev = None
        if self.pending:
            self.ev[s ^ 1].synchronize()
            prev = float(self.host[s ^ 1])
        self.pending = True
        self.slot = s ^ 1
        return prev

    def flush(self):
        if not self.pending:
            return None
        s = self.slot ^ 1
        self.ev[s].synchronize()
        self.pending = False
        return float(self.host[s])


# Task mix of the multitask loader (configs/learning_datasets/all.yaml concatenated and shuffled,
# datasets/coco_multitask_dataset.py:15-42), as SURVEY 8d "Config 3" fixes it for synthetic data:
# (task, probability, answer words lo..hi or None, boxes lo..hi or None).  Only detection samples carry boxes
# (datasets/coco_datasets.py:23-70,191); samples without an answer are encoded as `__cls__ __stop__` (gpv.py:404-406).
TASK_MIX = (("CocoCaptioning", 0.35, (8, 18), None), ("CocoVqa", 0.35, (1, 3), None),
            ("CocoDetection", 0.15, None, (1, 10)), ("CocoClassification", 0.15, (1, 1), None))


class SyntheticMultitask:
    """`n` synthetic batches shaped like the reference's multitask stream: normalised images [B,3,H,W], `Tl` query token
    ids per sample, and per-sample target dicts whose task is drawn from TASK_MIX.  The answer length -- hence the
    teacher-forced sequence length S = batch maximum + 2 -- varies from batch to batch, as in the reference."""

    def __init__(self, n, batch_size, H, W, vocab, seed=0, Tl=20):
        self.n, self.B, self.H, self.W, self.seed, self.Tl = n, batch_size, H, W, seed, Tl
        self.words = [w for w in vocab if not w.startswith("__")]

    def __len__(self):
        return self.n

    def targets(self, g):
        cum = torch.tensor([p for _, p, _, _ in TASK_MIX]).cumsum(0)
        out = []
        for _ in range(self.B):
            ti = min(int(torch.searchsorted(cum, torch.rand(1, generator=g))), len(TASK_MIX) - 1)
            task, _, words, boxes = TASK_MIX[ti]
            t = {"task": task}
            if words is not None:
                nw = int(torch.randint(words[0], words[1] + 1, (1,), generator=g))
                t["answer"] = " ".join(self.words[int(i)] for i in torch.randint(0, len(self.words), (nw,), generator=g))
            if boxes is not None:
                nb = int(torch.randint(boxes[0], boxes[1] + 1, (1,), generator=g))
                t["boxes"] = torch.cat((0.25 + 0.5 * torch.rand(nb, 2, generator=g), 0.05 + 0.3 * torch.rand(nb, 2, generator=g)), -1)
                t["labels"] = torch.zeros(nb, dtype=torch.long)
            out.append(t)
        return out

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.n):
            imgs = torch.randn(self.B, 3, self.H, self.W, generator=g)
            qids = torch.randint(1000, 30000, (self.B, self.Tl), generator=g)
            yield imgs, qids, self.targets(g)
