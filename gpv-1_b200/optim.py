"""Fused gradient clipping + AdamW for the GPV training step (reference: exp/gpv/train_distr.py:228-253, 414-428).

The reference clips the gradient norm of the `detr_backbone` + `detr_head` parameter groups to
`cfg.training.clip_max_norm` (0.1) with `torch.nn.utils.clip_grad_norm_` and then steps `torch.optim.AdamW` with four
groups (backbone at `lr_backbone` = 1e-5, everything else at `lr` = 1e-4, weight decay 1e-4).  Group membership is by
substring, in this order: 'detr.backbone' -> backbone; 'detr' -> head (this catches `detr_joiner`); 'bert.' -> bert
(which also catches `bert_joiner`); the rest -> others.

`ClipAdamW` does both in two kernel launches over the engine's flat gradient arena (csrc/optim.cu): one sum of
squares over the clipped subset, one multi-tensor update that applies the clip coefficient (writing the clipped
gradients back, as `clip_grad_norm_` does) and the AdamW arithmetic of torch's single-tensor implementation.  The
optimizer state (exp_avg, exp_avg_sq) lives in two arenas with the gradient arena's layout.  `state_dict()` writes torch.optim's
own layout -- `state` keyed by the parameter's index in the reference's four-group order (train_distr.py:234-253: every
named parameter, group by group), each with its own `step`, plus `param_groups` -- so that the reference's
`optimizer.load_state_dict(ckpt['optimizer'])` (train_distr.py:273) reads it, and `load_state_dict()` accepts that layout
(from either implementation) as well as this module's earlier name-keyed form.  Parameters that never receive a gradient
(BERT, frozen stem / layer1, `vision_token` ...) are skipped, as torch skips parameters whose `.grad` is None.  Known
deviation: a trainable parameter whose gradient is None in a particular step of the reference (e.g. the box heads on a
batch without any box target) is skipped there; here its arena slice is zero and it takes a zero-gradient step (weight
decay and moment decay only).
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
                 clip_max_norm=0.1, clip_groups=(0, 1), on_step=None, all_names=None):
        """named: list of (name, parameter tensor (fp32, CUDA, contiguous), gradient view inside `grad_arena`).
        all_names: every parameter name of the module in named_parameters() order (the reference hands all of them to AdamW);
        defaults to the names in `named`."""
        lib = _C.lib()
        self.arena = grad_arena
        self.names = [n for n, _, _ in named]
        self.params = [p for _, p, _ in named]
        self.lrs = [lr_backbone, lr, lr, lr]
        self.weight_decay, self.betas, self.eps, self.clip_max_norm = weight_decay, betas, eps, clip_max_norm
        self.t = 0
        self.step0 = [0] * len(named)                         # per tensor: the global step at which its state started
        self.on_step = on_step
        # torch.optim index of every parameter: group by group in named_parameters() order (train_distr.py:234-253)
        order = list(all_names) if all_names is not None else [n for n, _, _ in named]
        self.group_names = [[n for n in order if group_of(n) == g] for g in range(4)]
        self.index_of = {n: i for i, n in enumerate(n for grp in self.group_names for n in grp)}
        dev = grad_arena.device
        self.exp_avg = torch.zeros_like(grad_arena)
        self.exp_avg_sq = torch.zeros_like(grad_arena)
        self.total_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        chunk = lib.gpvb200_optim_chunk()
        rec = np.zeros(len(named), dtype=np.dtype([("p", "<u8"), ("goff", "<i8"), ("n", "<i4"), ("group", "<i4"), ("clip", "<i4"),
                                                   ("pad", "<i4")]))   # pad = step0
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
        self._rec = rec
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
        return cls(named, eng.grad_arena, on_step=eng.mark_dirty, all_names=list(params), **kw)

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

    def _set_step0(self, step0):
        self.step0 = [int(x) for x in step0]
        self._rec["pad"] = np.asarray(self.step0, dtype=np.int32)
        self.items.copy_(torch.from_numpy(self._rec.view(np.uint8).copy()))

    def state_dict(self):
        """torch.optim.AdamW.state_dict() layout (see the module docstring)."""
        state = {}
        for i, n in enumerate(self.names):
            off, num, shape = self.offsets[n]
            if self.t - self.step0[i] <= 0:
                continue                                       # never stepped: torch has no entry either
            state[self.index_of[n]] = {"step": torch.tensor(float(self.t - self.step0[i])),
                                       "exp_avg": self.exp_avg[off:off + num].view(shape).clone(),
                                       "exp_avg_sq": self.exp_avg_sq[off:off + num].view(shape).clone()}
        groups, base = [], 0
        for g, names in enumerate(self.group_names):
            groups.append({"lr": self.lrs[g], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                           "fused": None, "params": list(range(base, base + len(names)))})
            base += len(names)
        return {"state": state, "param_groups": groups, "clip_max_norm": self.clip_max_norm}

    def load_state_dict(self, sd):
        """Accepts torch.optim's layout (index-keyed `state` + `param_groups`, from the reference or from state_dict() above) or this
        module's earlier name-keyed form ({'state': {name: ...}, 't': ...}).  Tensors without an entry (e.g. DETR parameters that
        were frozen while the checkpoint was written) start from zero moments and step 0; entries for names that are not live here
        are ignored."""
        state = sd["state"]
        by_name = {}
        if "param_groups" in sd:
            flat = [i for g in sd["param_groups"] for i in g["params"]]
            names = [n for grp in self.group_names for n in grp]
            if len(flat) != len(names):
                raise ValueError(f"optimizer state has {len(flat)} parameters, this model has {len(names)}: not the same architecture")
            for pos, idx in enumerate(flat):
                if idx in state:
                    by_name[names[pos]] = state[idx]
        else:
            by_name = dict(state)
        steps = {n: int(float(s["step"])) if "step" in s else int(sd.get("t", 0)) for n, s in by_name.items()}
        self.t = max([int(sd.get("t", 0))] + list(steps.values()))
        step0 = []
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, n in enumerate(self.names):
            s = by_name.get(n)
            if s is None:
                step0.append(self.t)
                continue
            off, num, _ = self.offsets[n]
            self.exp_avg[off:off + num].copy_(s["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + num].copy_(s["exp_avg_sq"].reshape(-1))
            step0.append(self.t - steps[n])
        self._set_step0(step0)
