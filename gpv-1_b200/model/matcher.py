"""Stand-alone mirror of utils/matcher.py:HungarianMatcher on the device kernels (cost blocks + LSAP, no scipy)."""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as k


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou = float(cost_class), float(cost_bbox), float(cost_giou)

    @torch.no_grad()
    def forward(self, outputs, targets):
        """outputs: {'pred_relevance_logits' [B,Q,C], 'pred_boxes' [B,Q,4]} (CUDA); targets: list of {'labels' [T_b],
        'boxes' [T_b,4]}.  Returns [(idx_q, idx_t)] int64 CPU tensors like matcher.py:77."""
        logits = outputs["pred_relevance_logits"].float().contiguous()
        boxes = outputs["pred_boxes"].float().contiguous()
        dev = logits.device
        B, Q = logits.shape[:2]
        sizes = [int(t["boxes"].shape[0]) for t in targets]
        off = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(dev)
        Tmax = max(sizes) if sizes else 0
        if Tmax == 0:
            e = torch.empty(0, dtype=torch.int64)
            return [(e, e) for _ in range(B)]
        tb = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(dev, torch.float32)
        tl = torch.cat([t["labels"] for t in targets]).to(dev, torch.int64)
        cost = k.matcher_cost(logits, boxes, tb, tl, off, Tmax, self.cost_class, self.cost_bbox, self.cost_giou)
        iq, it = k.lsap(cost, off)
        iq, it = iq.cpu(), it.cpu()
        return [(iq[b, :min(Q, n)].clone(), it[b, :min(Q, n)].clone()) for b, n in enumerate(sizes)]
