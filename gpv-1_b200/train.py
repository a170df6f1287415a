"""Training entry point: the reference's train loop (exp/gpv/train_distr.py:151-470) on the B200 path.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m gpv1_b200.train training.batch_size=32 exp_dir=/tmp/exp

What is kept from the reference: the `training:` config block (configs/exp/gpv.yaml:119-152) with Hydra-style
`key=value` overrides, one process per GPU, `model.encode_answers` -> `model(imgs, queries, answer_token_ids, targets)`
-> backward -> clip_grad_norm_(detr_backbone + detr_head, clip_max_norm) -> AdamW with four groups
(train_distr.py:228-253, 410-428), WarmupLinearSchedule stepped every iteration when `lr_linear_decay` else
MultiStepLR(lr_milestones, lr_drop) per epoch (286-311, 466-470), the two-phase schedule (`training.freeze`:
freeze_detr_params for `frozen_epochs`, 136-140, 160-161, 322-323), checkpoints with `module.`-prefixed keys plus
optimizer / epoch / step (382-394) and resume from `training.ckpt` (262-285).

What is replaced: DistributedDataParallel(find_unused_parameters=True) by parallel.GradSync (stage-bucketed NCCL
all-reduce of the gradient arena, overlapped with backward) and clip + AdamW by optim.ClipAdamW (two launches); the synchronous
`imgs.to(gpu)` (401) by data.DevicePrefetcher (the next batch is copied on a copy stream under the current step) and the per-step
`loss.item()` of the log line (433) by data.LossReader (every loss reaches the host one step late, the GPU queue never drains).  The step
is replayed from CUDA graphs: the multitask stream changes shape from batch to batch (answers are padded to the batch maximum), so
the first `training.cuda_graphs` distinct batch shapes (image size, query length, answer length; default 4, each owns ~4 GB of saved
activations at B = 32) are captured the first time they are seen (`GPV.auto_capture` -> `GPV.capture_step(add=True)`) and replayed
afterwards; other shapes and mixed-size (padded) batches issue the step's ~1000 kernels eagerly.  A checkpoint is written at the end of every epoch when
`ckpt_dir` or `exp_dir` is set.  Data sets, evaluation and visualisation are out of scope
(SURVEY 8): `data` is any iterable of `(images, queries, targets)` batches as utils/detr_misc.py:collate_fn yields them;
without one, a synthetic loader of the reference's batch shape is used so that the entry point runs stand-alone.
"""
import math
import os
import sys
import time

import torch
import torch.distributed as dist

from .config import load_config
from .data import DevicePrefetcher, LossReader
from .model import GPV
from .optim import ClipAdamW
from .parallel import GradSync, broadcast_parameters


def lr_multiplier(step, total_steps, warmup_fraction):
    """pytorch_transformers.WarmupLinearSchedule (train_distr.py:296-302): linear 0 -> 1 over warmup_steps, then linear
    1 -> 0 at t_total."""
    warmup = warmup_fraction * total_steps
    if step < warmup:
        return float(step) / float(max(1.0, warmup))
    return max(0.0, float(total_steps - step) / float(max(1.0, total_steps - warmup)))


def multistep_multiplier(epoch, milestones, gamma):
    """torch.optim.lr_scheduler.MultiStepLR (train_distr.py:287-291), used when lr_linear_decay is off."""
    return gamma ** sum(1 for m in milestones if epoch >= m)


def freeze_detr_params(model, requires_grad=False):
    """train_distr.py:136-140: the parameters initialised from the pretrained DETR stay fixed in phase one."""
    init = set(model.init_detr_params)
    for n, p in model.named_parameters():
        if n in init:
            p.requires_grad = requires_grad


def save_checkpoint(path, model, optimizer, epoch, step, metric=0.0, lr_scale=1.0):
    """train_distr.py:382-394 layout: 'model' keys carry DDP's `module.` prefix so that the reference's inference.py:57-62
    and its own resume code read the file unchanged: `optimizer` is torch.optim.AdamW's state_dict layout (optim.ClipAdamW.state_dict),
    `lr` and `warmup_scheduler` are the entries train_distr.py:309-311 looks up."""
    sd = {f"module.{k}": v.detach().cpu() for k, v in model.state_dict().items()}
    osd = optimizer.state_dict()
    eng = getattr(model, "_engine", None)
    torch.save({"model": sd, "optimizer": osd, "epoch": epoch, "step": step, "model_selection_metric": metric,
                "drop_seed": int(eng.drop_seed.item()) & ((1 << 40) - 1) if eng is not None else 0,   # dropout step counter (rank bits stripped)
                "lr": [g["lr"] * lr_scale for g in osd.get("param_groups", [])],
                # WarmupLinearSchedule.state_dict() of the reference (LambdaLR): what its resume path reads back (train_distr.py:309-311)
                "warmup_scheduler": {"last_epoch": step, "_step_count": step + 1, "base_lrs": list(getattr(optimizer, "lrs", []))}}, path)


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """train_distr.py:262-285: same-shaped tensors only, with or without the `module.` prefix.  Returns (epoch, step)."""
    ckpt = torch.load(path, map_location=map_location)
    cur = model.state_dict()
    for k, v in ckpt["model"].items():
        k = k[len("module."):] if k.startswith("module.") else k
        if k in cur and cur[k].size() == v.size():
            cur[k] = v
    model.load_state_dict(cur)
    if "drop_seed" in ckpt and getattr(model, "_engine", None) is not None:        # resumed runs continue the mask sequence
        eng = model._engine
        eng.drop_seed.copy_((eng.drop_seed >> 40 << 40) + int(ckpt["drop_seed"]))
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])          # torch.optim layout (either implementation) or the earlier name-keyed form; raises on a mismatch
    return int(ckpt.get("epoch", -1)), int(ckpt.get("step", 0))


class SyntheticBatches:
    """Stand-in for the COCO loaders (out of scope): `n` batches of the reference's shapes -- normalised images
    [B,3,H,W], query token ids, multitask targets with boxes (cxcywh in (0,1)), labels 0 and an answer string."""

    def __init__(self, n, batch_size, H, W, vocab, seed=0, Tl=20):
        self.n, self.B, self.H, self.W, self.vocab, self.seed, self.Tl = n, batch_size, H, W, vocab, seed, Tl

    def __len__(self):
        return self.n

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        tasks = ["CocoCaptioning", "CocoVqa", "CocoDetection", "CocoClassification"]
        words = [w for w in self.vocab if not w.startswith("__")]
        for _ in range(self.n):
            imgs = torch.randn(self.B, 3, self.H, self.W, generator=g)
            qids = torch.randint(1000, 30000, (self.B, self.Tl), generator=g)
            targets = []
            for b in range(self.B):
                task = tasks[int(torch.randint(0, 4, (1,), generator=g))]
                t = {"task": task}
                if task != "CocoDetection":
                    nw = {"CocoCaptioning": 10, "CocoVqa": 2, "CocoClassification": 1}[task]
                    t["answer"] = " ".join(words[int(i)] for i in torch.randint(0, len(words), (nw,), generator=g))
                if task in ("CocoDetection", "CocoCaptioning"):
                    nb = int(torch.randint(1, 9, (1,), generator=g))
                    t["boxes"] = torch.cat((0.25 + 0.5 * torch.rand(nb, 2, generator=g), 0.05 + 0.3 * torch.rand(nb, 2, generator=g)), -1)
                    t["labels"] = torch.zeros(nb, dtype=torch.long)
                targets.append(t)
            yield imgs, qids, targets


def train(cfg, data=None, vocab=None, vocab_embed=None, log=print):
    """One process of the data-parallel job (train_distr.py:151-470 `train_worker`).  Returns the last loss value."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    tr = cfg.training
    model = GPV(cfg.model, vocab=vocab, vocab_embed=vocab_embed).to(dev)
    if getattr(cfg.model, "pretr_detr", None) and os.path.exists(str(cfg.model.pretr_detr)):
        model.load_pretr_detr()
    if tr.freeze is True:
        freeze_detr_params(model)
    broadcast_parameters(model)
    if world > 1:
        GradSync(model)
    model.auto_capture = int(getattr(tr, "cuda_graphs", 4))      # batch shapes replayed from CUDA graphs (0: eager launches only)
    optimizer = ClipAdamW.for_model(model, tr)
    last_epoch, step = -1, 0
    if tr.ckpt is not None:
        last_epoch, step = load_checkpoint(tr.ckpt, model, optimizer, map_location=dev)
        log(f"Loading checkpoint at the end of epoch {last_epoch}")
    epochs = int(tr.frozen_epochs if tr.freeze is True else tr.num_epochs)
    bs = int(tr.frozen_batch_size if tr.freeze is True else tr.batch_size) // world
    if data is None:
        data = SyntheticBatches(int(getattr(tr, "synthetic_iters", 8)), bs, 480, 640, model.vocab, seed=1000 + rank)
    total_steps = len(data) * epochs
    loss_val, prev_lr = None, 0.0
    reader = LossReader()
    for epoch in range(last_epoch + 1, epochs):
        # batches cross PCIe one step ahead on a copy stream (pinned staging buffers, device ping-pong buffers: data.DevicePrefetcher)
        for it, (imgs, queries, targets) in enumerate(DevicePrefetcher(iter(data), dev)):
            model.train()
            _, answer_token_ids = model.encode_answers(targets)
            for i, t in enumerate(targets):
                t["answer_token_ids"] = answer_token_ids[i, 1:]
            total_loss = model(imgs, queries, answer_token_ids, targets)
            if total_loss is not None:
                optimizer.zero_grad()
                total_loss.backward()
                if tr.lr_linear_decay and tr.lr_warmup:
                    mult = lr_multiplier(step + 1, total_steps, float(tr.lr_warmup_fraction))   # LambdaLR(last_epoch=step) has already stepped once
                else:
                    mult = multistep_multiplier(epoch, list(tr.lr_milestones), float(tr.lr_drop))
                optimizer.step(lr_scale=mult)
                # the loss of every step reaches the host one step late (data.LossReader): the log line never drains the GPU queue
                prev = reader.push(total_loss)
                if prev is not None:
                    loss_val = prev
                    if rank == 0 and (step - 1) % int(tr.log_step) == 0:
                        log(f"Epoch: {epoch} | Iter: {it} | Step: {step - 1} |  LR: {prev_lr:.3e} | total_loss: {round(loss_val, 4)}")
                prev_lr = optimizer.lrs[1] * mult
            step += 1
        ckpt_dir = getattr(cfg, "ckpt_dir", None) or (os.path.join(str(cfg.exp_dir), "ckpts") if getattr(cfg, "exp_dir", None) else None)
        if rank == 0 and ckpt_dir:
            os.makedirs(ckpt_dir, exist_ok=True)
            save_checkpoint(os.path.join(ckpt_dir, "model.pth"), model, optimizer, epoch, step, lr_scale=mult if total_loss is not None else 1.0)
    last = reader.flush()
    if last is not None:
        loss_val = last
        if rank == 0 and (step - 1) % int(tr.log_step) == 0:
            log(f"Step: {step - 1} |  LR: {prev_lr:.3e} | total_loss: {round(loss_val, 4)}")
    if world > 1:
        dist.barrier()
    return loss_val


def main(argv=None):
    cfg = load_config(overrides=list(sys.argv[1:] if argv is None else argv))
    t0 = time.time()
    train(cfg)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"done in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
