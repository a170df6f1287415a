"""Stand-alone mirrors of the reference's criterion modules on the device kernels, for call sites that evaluate the
losses of an output dictionary outside the fused training step (logging, validation):

    GPVCriterion(cfg.losses)(outputs, targets) -> (total_loss, loss_dict)         exp/gpv/models/losses.py:141-176
    SetCriterion(...)(outputs, targets)        -> {'loss_ce','loss_bbox','loss_giou'}   utils/set_criterion.py:150-191

Same filtering rules (answer losses per task over the samples that carry 'answer' and that task; localisation over
the samples that carry 'boxes'), same normalisers (mean over the filtered batch, sum over positions; class-weighted CE
with eos_coef; box losses / num_boxes with num_boxes rank-local), same weights.  VALUES ONLY: the tensors returned
here are not connected to autograd -- gradients of the training step come from GPV.forward(..., targets), whose fused
criterion kernels produce loss and gradient together.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as k
from .matcher import HungarianMatcher

TASK_LOSS = {"CocoCaptioning": "loss_caption", "CocoVqa": "loss_vqa", "CocoClassification": "loss_cls"}


class SetCriterion(nn.Module):
    def __init__(self, num_classes=1, matcher=None, weight_dict=None, eos_coef=0.1, losses=("labels", "boxes")):
        super().__init__()
        self.num_classes, self.matcher, self.weight_dict, self.eos_coef, self.losses = num_classes, matcher, weight_dict, float(eos_coef), list(losses)
        w = torch.ones(num_classes + 1)
        w[-1] = self.eos_coef
        self.register_buffer("empty_weight", w)

    @torch.no_grad()
    def forward(self, outputs, targets):
        logits = outputs["pred_relevance_logits"].float().contiguous()
        boxes = outputs["pred_boxes"].float().contiguous()
        dev = logits.device
        B, Q = logits.shape[:2]
        sizes = [int(t["boxes"].shape[0]) for t in targets]
        off = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(dev)
        sumT, Tmax = sum(sizes), max(sizes) if sizes else 0
        idx_q = idx_t = None
        tb = torch.zeros((1, 4), device=dev)
        if Tmax:
            tb = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(dev, torch.float32)
            tl = torch.cat([t["labels"] for t in targets]).to(dev, torch.int64)
            m = self.matcher
            cost = k.matcher_cost(logits, boxes, tb, tl, off, Tmax, m.cost_class, m.cost_bbox, m.cost_giou)
            idx_q, idx_t = k.lsap(cost, off)
        n_match = sum(min(Q, s) for s in sizes)
        weight_sum = float(n_match + self.eos_coef * (B * Q - n_match))
        out3 = torch.zeros(3, device=dev)
        lg8 = torch.zeros((B * Q, 8), device=dev)
        lg8[:, :logits.shape[2]] = logits.view(B * Q, -1)
        bx8 = torch.zeros((B * Q, 8), device=dev)
        bx8[:, :4] = boxes.view(B * Q, 4)
        k.set_criterion(lg8, bx8, tb, off, idx_q, idx_t, torch.ones(B, dtype=torch.uint8, device=dev), eos_coef=self.eos_coef,
                        weight_sum=weight_sum, num_boxes=float(max(sumT, 1)), wt_ce=1.0, wt_bbox=1.0, wt_giou=1.0, out3=out3,
                        dlogits=torch.zeros((B * Q, 8), device=dev), dbox_pre=torch.zeros((B * Q, 8), device=dev, dtype=torch.bfloat16))
        return {"loss_ce": out3[0], "loss_bbox": out3[1], "loss_giou": out3[2]}


class GPVCriterion(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.loss_wts = {}
        for _, lc in cfg.items():
            self.loss_wts.update({kk: float(v) for kk, v in lc.loss_wts.items()})
        loc = cfg.Localization
        self.matcher = HungarianMatcher(cost_class=loc.cost_wts.ce, cost_bbox=loc.cost_wts.bbox, cost_giou=loc.cost_wts.giou)
        self.set_criterion = SetCriterion(num_classes=loc.num_classes, matcher=self.matcher, eos_coef=loc.eos_coef)

    @torch.no_grad()
    def forward(self, outputs, targets):
        loss_dict = {}
        lg = outputs["answer_logits"][-1].float().contiguous()          # [B, S', V] of the last decoder layer (L = 1)
        B, S1, V = lg.shape
        dev = lg.device
        for task, name in TASK_LOSS.items():
            if name not in self.loss_wts:
                continue
            rows = [b for b, t in enumerate(targets) if "answer" in t and t.get("task") == task]
            if not rows:
                loss_dict[name] = None
                continue
            w = torch.zeros((B, S1), device=dev)
            w[rows] = 1.0 / len(rows)                                     # losses.py:26: mean over the filtered batch, sum over positions
            tg = torch.zeros((B, S1), dtype=torch.int64, device=dev)
            for b in rows:
                tg[b] = targets[b]["answer_token_ids"].to(dev)
            out = torch.zeros(1, device=dev)
            k.ce_fwd_bwd(lg.view(B * S1, V), tg.view(-1), w.view(-1), out, None)
            loss_dict[name] = out[0]
        idxs = [b for b, t in enumerate(targets) if "boxes" in t]
        if idxs:
            sub = {"pred_relevance_logits": outputs["pred_relevance_logits"][idxs], "pred_boxes": outputs["pred_boxes"][idxs]}
            loss_dict.update(self.set_criterion(sub, [targets[b] for b in idxs]))
        else:
            loss_dict.update({"loss_ce": None, "loss_bbox": None, "loss_giou": None})
        if all(v is None for v in loss_dict.values()):
            return None, loss_dict
        total = 0
        for name, wt in self.loss_wts.items():
            if loss_dict.get(name) is not None:
                total = total + wt * loss_dict[name]
        return total, loss_dict
