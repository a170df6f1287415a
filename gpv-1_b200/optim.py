"""Fused gradient clipping + AdamW for the GPV training step (reference: exp/gpv/train_distr.py:228-253, 414-428).

The reference clips the gradient norm of the `detr_backbone` + `detr_head` parameter groups to
`cfg.training.clip_max_norm` (0.1) with `torch.nn.utils.clip_grad_norm_` and then steps `torch.optim.AdamW` with four
groups (backbone at `lr_backbone` = 1e-5, everything else at `lr` = 1e-4, weight decay 1e-4).  Group membership is by
substring, in this order: 'detr.backbone' -> backbone; 'detr' -> head (this catches `detr_joiner`); 'bert.' -> bert
(which also catches `bert_joiner`); the rest -> others.

`ClipAdamW` does both in two kernel launches over the engine's flat gradient arena (csrc/optim.cu): one sum of
squares over the clipped subset, one multi-tensor update that applies the clip coefficient (writing the clipped
gradients back, as `clip_grad_norm_` does) and the AdamW arithmetic of torch's single-tensor implementation.  The
optimizer state (exp_avg, exp_avg_sq) lives in two arenas with the gradient arena's layout; `state_dict()` /
`load_state_dict()` expose it per parameter name with torch.optim's key names so checkpoints stay interchangeable.
Parameters that never receive a gradient (BERT, frozen stem / layer1, `vision_token` ...) are skipped, as torch skips
parameters whose `.grad` is None.
"""
import ctypes

import numpy as np
import torch

from . import _C

GROUPS = ("detr_backbone", "detr_head", "bert", "others")


def group_of(name: str) -> int:
    """train_distr.py:234-242."""
    if "detr.backbone" in name:
        return 0
    if "detr" in name:
        return 1
    if "bert." in name:
        return 2
    return 3


class ClipAdamW:
    def __init__(self, named, grad_arena, *, lr=1e-4, lr_backbone=1e-5, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 clip_max_norm=0.1, clip_groups=(0, 1), on_step=None):
        """named: list of (name, parameter tensor (fp32, CUDA, contiguous), gradient view inside `grad_arena`)."""
        lib = _C.lib()
        self.arena = grad_arena
        self.names = [n for n, _, _ in named]
        self.params = [p for _, p, _ in named]
        self.lrs = [lr_backbone, lr, lr, lr]
        self.weight_decay, self.betas, self.eps, self.clip_max_norm = weight_decay, betas, eps, clip_max_norm
        self.t = 0
        self.on_step = on_step
        dev = grad_arena.device
        self.exp_avg = torch.zeros_like(grad_arena)
        self.exp_avg_sq = torch.zeros_like(grad_arena)
        self.total_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        chunk = lib.gpvb200_optim_chunk()
        rec = np.zeros(len(named), dtype=np.dtype([("p", "<u8"), ("goff", "<i8"), ("n", "<i4"), ("group", "<i4"), ("clip", "<i4"),
                                                   ("pad", "<i4")]))
        assert rec.dtype.itemsize == lib.gpvb200_optim_item_size()
        bi, bc, ci, cc = [], [], [], []
        self.offsets = {}
        for i, (name, p, g) in enumerate(named):
            assert p.dtype == torch.float32 and p.is_cuda and p.is_contiguous(), name
            assert g.dtype == torch.float32 and g.is_contiguous() and g.numel() == p.numel(), name
            off = (g.data_ptr() - grad_arena.data_ptr()) // 4
            assert 0 <= off and off + p.numel() <= grad_arena.numel(), name
            grp = group_of(name)
            clip = int(grp in clip_groups)
            rec[i] = (p.data_ptr(), off, p.numel(), grp, clip, 0)
            self.offsets[name] = (off, p.numel(), tuple(p.shape))
            nb = (p.numel() + chunk - 1) // chunk
            bi += [i] * nb
            bc += list(range(nb))
            if clip:
                ci += [i] * nb
                cc += list(range(nb))
        self.items = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
        mk = lambda a: torch.tensor(a, dtype=torch.int32).to(dev)
        self.blk_item, self.blk_chunk, self.clip_item, self.clip_chunk = mk(bi), mk(bc), mk(ci), mk(cc)

    @classmethod
    def for_model(cls, model, training_cfg=None, **kw):
        """Optimizer over every parameter of a gpv1_b200 GPV module that can receive a gradient, with the reference's
        hyper-parameters (configs/exp/gpv.yaml:130-144) unless overridden."""
        eng = model.engine
        params = dict(model.named_parameters())
        named = [(n, params[n].data, eng.G[n]) for n in eng.live_names if params[n].requires_grad]   # torch skips params without grad
        if training_cfg is not None:
            for key in ("lr", "lr_backbone", "weight_decay", "clip_max_norm"):
                if key not in kw and hasattr(training_cfg, key):
                    kw[key] = float(getattr(training_cfg, key))
        return cls(named, eng.grad_arena, on_step=eng.mark_dirty, **kw)

    @torch.no_grad()
    def step(self, lr_scale=1.0):
        """One optimizer step on the gradients currently in the arena.  lr_scale: the scheduler's multiplier
        (WarmupLinearSchedule x MultiStepLR in the reference, train_distr.py:286-311)."""
        lib = _C.lib()
        self.t += 1
        st = _C.stream_ptr()
        f = ctypes.c_float
        _C.check(lib.gpvb200_grad_sqnorm(_C.ptr(self.items), _C.ptr(self.clip_item), _C.ptr(self.clip_chunk), self.clip_item.numel(),
                                         _C.ptr(self.arena), _C.ptr(self.total_sq), st), "grad_sqnorm")
        lr = [x * lr_scale for x in self.lrs]
        _C.check(lib.gpvb200_clip_adamw(_C.ptr(self.items), _C.ptr(self.blk_item), _C.ptr(self.blk_chunk), self.blk_item.numel(),
                                        _C.ptr(self.arena), _C.ptr(self.exp_avg), _C.ptr(self.exp_avg_sq), _C.ptr(self.total_sq),
                                        f(self.clip_max_norm), f(lr[0]), f(lr[1]), f(lr[2]), f(lr[3]), f(self.betas[0]),
                                        f(self.betas[1]), f(self.eps), f(self.weight_decay), ctypes.c_int64(self.t), st), "clip_adamw")
        if self.on_step is not None:
            self.on_step()        # the engine's packed bf16 weights are stale now

    def grad_norm(self):
        """Total norm of the clipped subset as seen by the last step() (what clip_grad_norm_ returns)."""
        return self.total_sq.sqrt()

    def zero_grad(self, set_to_none=True):
        pass                      # the engine zeroes the arena at the start of every backward

    def state_dict(self):
        st = {}
        for n, (off, num, shape) in self.offsets.items():
            st[n] = {"step": self.t, "exp_avg": self.exp_avg[off:off + num].view(shape).clone(),
                     "exp_avg_sq": self.exp_avg_sq[off:off + num].view(shape).clone()}
        return {"state": st, "t": self.t, "lrs": list(self.lrs), "weight_decay": self.weight_decay, "betas": self.betas, "eps": self.eps,
                "clip_max_norm": self.clip_max_norm}

    def load_state_dict(self, sd):
        self.t = int(sd["t"])
        for n, s in sd["state"].items():
            off, num, _ = self.offsets[n]
            self.exp_avg[off:off + num].copy_(s["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + num].copy_(s["exp_avg_sq"].reshape(-1))
