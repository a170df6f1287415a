"""Host-side mirror of the reference module surface for the hot path: `GPV` (exp/gpv/models/gpv.py:58-466).

Same constructor config (`cfg.model` of configs/exp/gpv.yaml), same `state_dict()` keys and shapes (836 entries, so
reference checkpoints load by name: inference.py:57-62, train_distr.py:264-272), same call signatures and return
values:

    loss = model(images, queries, answer_token_ids, targets)        # scalar; loss.backward() fills p.grad
    out  = model(images, queries, None)                             # greedy decode (gpv.py:178-196)
    out  = model.forward_beam_search(images, queries, beam_size)    # gpv.py:209-254

All arithmetic runs in the sm_100a kernels through `Engine`; there is no PyTorch-op or CPU fallback -- constructing
the engine on a machine without the built library or without a Blackwell GPU raises.
"""
import json
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import kernels as k
from .engine import Engine, TASK_LOSS
from .spec import gpv_specs, never_gets_grad

SPECIAL = ("__pad__", "__cls__", "__stop__", "__unk__")


def positionalencoding1d(d_model, length):
    """gpv.py:18-34 (sin/cos interleaved); only materialised as the frozen `pos_enc` parameter."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe


class _Node(nn.Module):
    """Bare container so that parameters keep the reference's dotted names (e.g. detr.backbone.0.body.conv1.weight)."""


def _register(root, dotted, tensor, kind):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if kind == "buffer":
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=(kind == "param")))


def _init_tensor(name, shape, g):
    """Random init in the spirit of the reference constructors (xavier/kaiming-scale matrices, unit norms, identity BN)."""
    if name.endswith(("running_var",)):
        return torch.ones(shape)
    if name.endswith(("running_mean",)):
        return torch.zeros(shape)
    is_norm = ".bn" in name or ".downsample.1." in name or "LayerNorm" in name or ".norm" in name
    if is_norm:
        return torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
    if name == "criterion.localization_criterion.set_criterion.empty_weight":
        return torch.ones(shape)
    if name == "pos_enc":
        return positionalencoding1d(shape[2], shape[1]).view(shape)
    if len(shape) == 4:
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
    if len(shape) == 2 and name not in ("relevance_tokens",):
        if name.endswith("_embeddings.weight"):
            return 0.02 * torch.randn(shape, generator=g)
        if name == "detr.query_embed.weight":
            return torch.randn(shape, generator=g)
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] + shape[1]))
    if name in ("vision_token", "lang_token", "relevance_tokens"):
        return 0.1 * torch.randn(shape, generator=g)
    return torch.zeros(shape)


class HostTargets:
    """Per-step host bookkeeping of the criterion (losses.py:41-138, set_criterion.py:150-168): which rows carry which
    loss weight, ragged target offsets, normalisers.  Built from Python-side shapes only -- no device sync.

    With `static` (a dict of preallocated device buffers from `alloc_static`) the results are copied into those fixed
    addresses (CUDA-graph replay) and the normalisers are left to the device (weight_sum = num_boxes = -1).  Host-resident
    targets (what a loader hands over) are packed on the host and cross PCIe as ONE copy; targets already on the device are
    gathered there piece by piece (`fast=False` forces that path for host tensors too: the reference for tests)."""

    def __init__(self, targets, B, S, Q, loss_wts, eos_coef, device, static=None, fast=True):
        self.device = device
        # ---- text losses: CE(reduction none).mean(0).sum(0).sum() per task (losses.py:20-26) -> weight wt/B' per row
        counts = {}
        for t in targets:
            if "answer" in t and t.get("task") in TASK_LOSS:
                counts[t["task"]] = counts.get(t["task"], 0) + 1
        roww = np.zeros((B, S), np.float32)
        answers = [None] * B                           # answer_token_ids of the rows that are supervised
        for b, t in enumerate(targets):
            if "answer" in t and t.get("task") in TASK_LOSS:
                roww[b, :S - 1] = loss_wts[TASK_LOSS[t["task"]]] / counts[t["task"]]
                ids = t["answer_token_ids"]
                if ids.shape[0] != S - 1:
                    raise ValueError("targets[i]['answer_token_ids'] must be answer_token_ids[i, 1:] (train_distr.py:410-412)")
                answers[b] = ids
        self.n_text = sum(counts.values())
        # ---- localisation: images that have a 'boxes' key (losses.py:101-121); T_b may be 0
        sizes = [int(t["boxes"].shape[0]) if "boxes" in t else 0 for t in targets]
        valid = [1 if "boxes" in t else 0 for t in targets]
        self.n_loc = sum(valid)
        self.Tmax = max(sizes) if sizes else 0
        sumT = sum(sizes)
        n_match = sum(min(Q, s) for s in sizes)
        self.weight_sum = float(n_match + eos_coef * (self.n_loc * Q - n_match)) if self.n_loc else 1.0
        self.num_boxes = float(max(sumT, 1))
        self.sizes = sizes
        host = np.zeros(B + 1 + B, np.int32)
        host[1:B + 1] = np.cumsum(sizes)
        host[B + 1:] = valid
        head = np.concatenate([host.view(np.float32), roww.reshape(-1)])      # offsets | loc_valid | CE row weights
        with_boxes = [t for t in targets if "boxes" in t and t["boxes"].shape[0]]
        on_host = all(not v.is_cuda for t in targets for kk, v in t.items()
                      if kk in ("boxes", "labels", "answer_token_ids") and torch.is_tensor(v))

        def pinned(a):
            h = torch.from_numpy(a)
            return h.pin_memory() if torch.device(device).type == "cuda" else h

        def ce_rows():                                  # [B, S-1] int64 on the device, zero rows where nothing is supervised
            zero = torch.zeros(S - 1, dtype=torch.int64, device=device)
            return torch.stack([zero if a is None else a.to(device=device, dtype=torch.int64, non_blocking=True) for a in answers])

        def cat_boxes():
            return (torch.cat([t["boxes"].to(device=device, dtype=torch.float32, non_blocking=True).reshape(-1, 4) for t in with_boxes]),
                    torch.cat([t["labels"].to(device=device, dtype=torch.int64, non_blocking=True) for t in with_boxes]))

        if static is None:
            dv = pinned(head).to(device, non_blocking=True)
            tg = torch.zeros((B, S), dtype=torch.int64, device=device)
            if S > 1:
                tg[:, :S - 1] = ce_rows()
            self.ce_targets = tg.view(-1)
            if sumT:
                self.boxes, self.labels = cat_boxes()
            else:
                self.boxes = torch.zeros((1, 4), device=device)
                self.labels = torch.zeros(1, dtype=torch.int64, device=device)
        else:
            if static["ce_targets"].numel() != B * S or self.Tmax > static["Tcap"]:
                raise ValueError("batch does not fit the captured step (B, S or boxes per image)")
            dv = static["packed"]
            staged = False
            if fast and on_host and not static.get("stage_failed"):
                try:
                    self._stage_host(static, head, answers, with_boxes, sumT, B, S)
                    staged = True
                except (RuntimeError, TypeError, ValueError) as e:      # e.g. no pinned memory left: the piecewise device path still works
                    static["stage_failed"] = repr(e)
            if not staged:
                dv.copy_(pinned(head), non_blocking=True)
                if S > 1:
                    static["ce_targets"].view(B, S)[:, :S - 1] = ce_rows()
                if sumT:
                    static["boxes"][:sumT], static["labels"][:sumT] = cat_boxes()
            self.ce_targets = static["ce_targets"]
            self.boxes, self.labels = static["boxes"], static["labels"]
            self.Tmax = static["Tcap"]
            self.n_loc = max(self.n_loc, 1)            # the graph always runs the localisation kernels
            self.weight_sum = self.num_boxes = -1.0     # derived on the device from offsets / loc_valid
        self.offsets = dv[:B + 1].view(torch.int32)
        self.loc_valid_i32 = dv[B + 1:2 * B + 1].view(torch.int32)
        if static is None:
            self.loc_valid = self.loc_valid_i32.to(torch.uint8)
        else:
            static["loc_valid"].copy_(self.loc_valid_i32)
            self.loc_valid = static["loc_valid"]
        self.ce_row_weight = dv[2 * B + 1:]

    @staticmethod
    def _stage_host(static, head, answers, with_boxes, sumT, B, S):
        """Pack head | CE targets | boxes | labels into the next of two pinned staging buffers laid out like the device blob
        and enqueue ONE host-to-device copy.  A staging buffer is rewritten only after the copy that last read it completed."""
        lay = static["layout"]
        turn = static["turn"] = (static["turn"] + 1) % len(static["stage"])
        ev = static["stage_ev"][turn]
        if ev is not None:
            ev.synchronize()
        stage = static["stage"][turn]
        buf = stage.numpy()
        o, n = lay["packed"]
        buf[o:o + n].view(np.float32)[:] = head
        o, n = lay["ce_targets"]
        ce = buf[o:o + n].view(np.int64).reshape(B, S)
        ce[:] = 0
        for b, a in enumerate(answers):
            if a is not None:
                ce[b, :S - 1] = a.numpy()
        if sumT:
            o, n = lay["boxes"]
            buf[o:o + n].view(np.float32).reshape(-1, 4)[:sumT] = np.concatenate([t["boxes"].numpy().reshape(-1, 4) for t in with_boxes])
            o, n = lay["labels"]
            buf[o:o + n].view(np.int64)[:sumT] = np.concatenate([t["labels"].numpy().reshape(-1) for t in with_boxes])
        static["blob"].copy_(stage, non_blocking=True)
        if static["blob"].is_cuda:
            ev = static["stage_ev"][turn] = ev if ev is not None else torch.cuda.Event()
            ev.record()

    @staticmethod
    def alloc_static(B, S, Tcap, device):
        """Fixed-address buffers of one captured step: ONE device blob with typed views (so that host-resident targets arrive
        in a single copy) and two pinned staging buffers of the same layout."""
        sizes = [("packed", (2 * B + 1 + B * S) * 4), ("ce_targets", B * S * 8), ("boxes", B * Tcap * 16), ("labels", B * Tcap * 8)]
        layout, off = {}, 0
        for name, n in sizes:
            layout[name] = (off, n)
            off += (n + 15) // 16 * 16
        blob = torch.zeros(off, dtype=torch.uint8, device=device)
        cuda = torch.device(device).type == "cuda"

        def view(name, dtype):
            o, n = layout[name]
            return blob[o:o + n].view(dtype)

        try:
            stage, failed = [torch.zeros(off, dtype=torch.uint8, pin_memory=cuda) for _ in range(2)], None
        except RuntimeError as e:                       # no pinned memory: targets take the piecewise path
            stage, failed = [], repr(e)
        return {"Tcap": Tcap, "blob": blob, "layout": layout, "turn": 0, "stage_ev": [None, None], "stage": stage, "stage_failed": failed,
                "packed": view("packed", torch.float32), "ce_targets": view("ce_targets", torch.int64),
                "boxes": view("boxes", torch.float32).view(B * Tcap, 4), "labels": view("labels", torch.int64),
                "loc_valid": torch.zeros(B, dtype=torch.uint8, device=device)}


class _Step(torch.autograd.Function):
    """One autograd node for the whole step: forward already ran in the engine; backward runs the engine's explicit
    backward and hands each parameter its slice of the gradient arena."""

    @staticmethod
    def forward(ctx, anchor, loss, model):
        ctx.model = model
        return loss.clone().view(())

    @staticmethod
    def backward(ctx, g):
        ctx.model._run_backward(g)
        return None, None, None


class GPV(nn.Module):
    def __init__(self, cfg, vocab=None, vocab_embed=None, seed=None):
        """cfg: the `model:` block (attribute access).  vocab / vocab_embed override cfg.vocab / cfg.vocab_embed paths
        (answer_head.py:61-74) when given as a list / array."""
        super().__init__()
        self.cfg = cfg
        if vocab is None:
            with open(cfg.vocab) as f:
                vocab = json.load(f)
        if vocab_embed is None and cfg.vocab_embed is not None:
            vocab_embed = np.load(cfg.vocab_embed)
        self.vocab = list(vocab)
        self.word_to_idx = {w: i for i, w in enumerate(self.vocab)}
        for w in SPECIAL:
            if w not in self.word_to_idx:
                raise ValueError(f"vocab lacks {w}")
        V = len(self.vocab)
        if cfg.roi_head is not True or cfg.answer_head == "linear" or cfg.detr.aux_loss or cfg.detr.pre_norm:
            raise NotImplementedError("only the default GPV-1 variant (roi_head, generated answer head, post-norm, no aux loss) is built")
        if getattr(cfg.text_decoder, "pos_enc", False):
            raise NotImplementedError("text_decoder.pos_enc=True (gpv.py:452 adds pos_enc to the answer embeddings) is not built: the "
                                      "shipped configuration keeps it False")
        for name in ("CocoVqa", "CocoClassification", "CocoCaptioning"):
            lc = getattr(cfg.losses, name, None)
            if lc is not None and getattr(lc, "pad_idx", None) is not None:
                raise NotImplementedError(f"losses.{name}.pad_idx (the cross-entropy ignore_index of losses.py:12-18) is not built: the "
                                          "shipped configuration keeps it null, every answer position carries weight")
        self.specs = gpv_specs(cfg, V)
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        for s in self.specs:
            t = _init_tensor(s.name, tuple(s.shape), g)
            _register(self, s.name, t, s.kind)
        if vocab_embed is not None:
            ve = torch.as_tensor(np.asarray(vocab_embed), dtype=torch.float32)
            if tuple(ve.shape) != (V, cfg.bert_joiner.bert_dim):
                raise ValueError(f"vocab_embed must be [{V},{cfg.bert_joiner.bert_dim}]")
        else:
            ve = 0.1 * torch.randn(V, cfg.bert_joiner.bert_dim, generator=g)
        with torch.no_grad():
            self.answer_head.vocab_embed.copy_(ve)
            self.answer_input_embedings.embedding_layer.weight.copy_(ve)
            self.criterion.localization_criterion.set_criterion.empty_weight[-1] = cfg.losses.Localization.eos_coef
        self.init_detr_params = []
        self._engine = None
        self._engine_key = None
        self._anchor_t = None
        self._tokenizer = None
        self.grad_sync = None          # parallel.GradSync installs itself here
        self._captured = None          # model/graph.py:CapturedStep once capture_step() ran (the one forward() tries first)
        self._captures = []            # every CapturedStep kept by capture_step(..., add=True): one per batch shape
        self.auto_capture = 0          # forward() captures up to this many batch shapes by itself (0: only explicit capture_step calls)
        self.inference_graphs = False  # True: greedy / beam inference of a repeated input shape replays one CUDA graph
        self._inf_graphs = {}

    # ------------------------------------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        dev = self.vision_token.device
        if dev.type != "cuda":
            raise RuntimeError("GPV (gpv-1_b200) runs only on a CUDA sm_100a device: move the module with .cuda() first; "
                               "there is no CPU path")
        key = (dev, self.vision_token.data_ptr())
        if self._engine is None or self._engine_key != key:
            tensors = {n: t for n, t in self.state_dict(keep_vars=True).items()}
            self._engine = Engine(tensors, self.specs, self.cfg, dev)
            self._engine_key = key
            self._live = [(n, tensors[n]) for n in self._engine.live_names]
            if self.grad_sync is not None:
                self.grad_sync.attach(self._engine)
        return self._engine

    def _anchor(self):
        dev = self.vision_token.device
        if self._anchor_t is None or self._anchor_t.device != dev:
            self._anchor_t = torch.zeros(1, device=dev, requires_grad=True)
        return self._anchor_t

    def capture_step(self, images, queries, answer_token_ids, targets, boxes_per_image_cap=None, add=False):
        """Record the training step for this batch shape into CUDA graphs (model/graph.py).  Later calls of
        forward(images, queries, answer_token_ids, targets) with the same shapes replay them.  `add=True` keeps the
        shapes captured before (each owns its graphs and memory pool): a multitask stream pads answers to the batch
        maximum (gpv.py:401-430), so the answer length S changes from step to step -- capture one step per S."""
        from .graph import CapturedStep
        self.sync_trainable()
        (images, mask), qids = self._images(images), self._queries(queries)
        if mask is not None:
            raise NotImplementedError("capture_step needs an unpadded batch (one image size): the padding mask is per batch")
        ans = answer_token_ids.to(device=images.device, dtype=torch.int64)
        self.engine.train_mode = bool(self.training)
        self._captured = CapturedStep(self, images, qids, ans, targets, boxes_per_image_cap)
        self._captures = (self._captures if add else []) + [self._captured]
        return self._captured

    def _run_backward(self, g):
        eng = self.engine
        cap = self._captured
        if cap is not None and cap.pending:
            cap.backward()
        else:
            eng.backward()
        if not eng.unit_upstream_grad:
            eng.grad_arena.mul_(g)
        for n, p in self._live:
            if not p.requires_grad:                   # frozen (train_distr.py:136-140 freeze_detr_params): autograd gives no grad
                p.grad = None
                continue
            gv = eng.G[n]
            if p.grad is None or p.grad.data_ptr() != gv.data_ptr():
                p.grad = gv

    def sync_trainable(self):
        """Tell the engine which parameters are frozen (`requires_grad = False`, e.g. after the reference's
        freeze_detr_params for the first training phase, train_distr.py:136-140,160-161): their weight- and bias-gradient
        kernels are skipped.  Called from forward; cheap (a tuple compare) when nothing changed."""
        eng = self.engine
        key = tuple(p.requires_grad for _, p in self._live)
        if key != getattr(self, "_trainable_key", None):
            eng.frozen = {n for n, p in self._live if not p.requires_grad}
            self._trainable_key = key
            if self._captured is not None:
                self._captured, self._captures = None, []   # the captured backward contains the old set of gradient kernels

    def load_pretr_detr(self):
        """gpv.py:122-135: copy same-shaped tensors of a DETR checkpoint under the `detr.` prefix."""
        loaded = torch.load(self.cfg.pretr_detr, map_location="cpu")["model"]
        cur = self.state_dict()
        for lk, v in loaded.items():
            dk = "detr." + lk
            if dk in cur and cur[dk].size() == v.size():
                self.init_detr_params.append(dk)
                cur[dk] = v
        self.load_state_dict(cur)

    # ------------------------------------------------------------------------------------------------ inputs
    def _images(self, images):
        """-> (tensor [B,3,H,W] fp32 or [B,H,W,3] uint8 on the device, padding mask [B,H,W] bool or None).
        Accepts what DETR.forward accepts (detr_roi_head.py:58-61): a NestedTensor, a list of [3,H,W] images (padded to
        the batch maximum with zeros, mask True on padding: utils/detr_misc.py:282-299) or a batched tensor."""
        dev = self.vision_token.device
        mask = None
        if hasattr(images, "tensors") and hasattr(images, "mask"):
            if images.mask is not None and bool(images.mask.any()):
                mask = images.mask.to(device=dev, dtype=torch.bool, non_blocking=True)
            images = images.tensors
        if isinstance(images, (list, tuple)):
            if len({tuple(i.shape) for i in images}) != 1:
                if images[0].dtype == torch.uint8:
                    raise NotImplementedError("mixed-size uint8 batches: pad on the loader side")
                H, W = max(i.shape[1] for i in images), max(i.shape[2] for i in images)
                batch = torch.zeros((len(images), 3, H, W), dtype=torch.float32)
                m = torch.ones((len(images), H, W), dtype=torch.bool)
                for b, im in enumerate(images):
                    batch[b, :, :im.shape[1], :im.shape[2]].copy_(im)
                    m[b, :im.shape[1], :im.shape[2]] = False
                images, mask = batch, m.to(dev, non_blocking=True)
            else:
                images = torch.stack(list(images))
        if images.dtype == torch.uint8:
            # raw loader format [B,H,W,3]: ToTensor + Normalize (coco_generic_dataset.py:31-32) are folded into the stem's read
            if images.dim() != 4 or images.shape[-1] != 3:
                raise ValueError("uint8 images must be NHWC [B,H,W,3]")
            return images.to(device=dev, non_blocking=True).contiguous(), mask
        return images.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(), mask

    def _queries(self, queries):
        dev = self.vision_token.device
        if torch.is_tensor(queries):
            return queries.to(device=dev, dtype=torch.int64, non_blocking=True)
        if len(queries) and not isinstance(queries[0], str):
            return torch.as_tensor(np.asarray(queries), dtype=torch.int64).to(dev, non_blocking=True)
        if self._tokenizer is None:
            from .tokenizer import load_tokenizer
            self._tokenizer = load_tokenizer(getattr(self.cfg, "bert_vocab", None))
        ids = self._tokenizer(list(queries))
        return torch.as_tensor(ids, dtype=torch.int64).to(dev, non_blocking=True)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, images, queries, answer_token_ids, targets=None, vocab_mask=None):
        eng = self.engine
        self.sync_trainable()
        # nn.Dropout semantics: masks in train(), identity in eval().  Generation (no answer_token_ids) always runs the
        # eval arithmetic: the reference only generates under model.eval() (inference.py:63, metrics.py)
        eng.train_mode = bool(self.training) and answer_token_ids is not None
        (images, mask), qids = self._images(images), self._queries(queries)
        B, Q = images.shape[0], self.cfg.detr.num_queries
        if answer_token_ids is not None and targets is not None:
            ans = answer_token_ids.to(device=images.device, dtype=torch.int64, non_blocking=True)
            S = ans.shape[1]
            cap = self._captured
            if cap is not None and mask is None and torch.is_grad_enabled() and not cap.matches(images, qids, ans):
                cap = next((c for c in self._captures if c.matches(images, qids, ans)), None)
                if cap is not None:
                    self._captured = cap              # _run_backward replays the backward graphs of the step that ran forward
            if (cap is None and self.auto_capture > len(self._captures) and mask is None and torch.is_grad_enabled()
                    and all(not c.matches(images, qids, ans) for c in self._captures)):
                # a batch shape seen for the first time (train.py sets auto_capture): record it, then replay it below and from now on
                cap = self.capture_step(images, qids, ans, targets, add=True)
                eng.train_mode = bool(self.training)
            if cap is not None and mask is None and torch.is_grad_enabled() and cap.matches(images, qids, ans):
                loss = cap.forward(images, qids, ans, targets)
                return None if loss is None else _Step.apply(self._anchor(), loss, self)
            tgt = HostTargets(targets, B, S, Q, eng.loss_wts, eng.eos_coef, images.device)
            if tgt.n_text == 0 and tgt.n_loc == 0:
                return None
            loss, _ = eng.forward_train(images, qids, ans, tgt, mask=mask)
            if torch.is_grad_enabled():
                return _Step.apply(self._anchor(), loss, self)
            eng.saved = None
            return loss.view(())
        if answer_token_ids is None and targets is None and self.inference_graphs and mask is None:
            return self._graphed("greedy", images, qids, vocab_mask, None)
        with torch.no_grad():
            s = eng.encode(images, qids, save=False, mask=mask)
            outputs = self._outputs(s)
            if answer_token_ids is None:
                outputs["answer_logits"] = self._greedy(s, vocab_mask)
            else:
                ans = answer_token_ids.to(device=images.device, dtype=torch.int64)
                S = ans.shape[1]
                lg, _ = eng.decode_text(ans.reshape(-1), s["memory"], B, S, s["Tm"], save=False)
                outputs["answer_logits"] = lg.view(1, B, S, -1)[:, :, :-1, :eng.V]
            if targets is not None:
                raise ValueError("targets given without answer_token_ids")
        return outputs

    def _outputs(self, s):
        B, Q = s["B"], self.cfg.detr.num_queries
        return {"pred_relevance_logits": s["logits"].view(B, Q, 8)[:, :, :2], "pred_boxes": s["boxes"].view(B, Q, 8)[:, :, :4],
                "detr_hs": s["detr_hs_joined"].float().view(1, B, Q, -1)}

    def _greedy(self, s, vocab_mask):
        """gpv.py:178-196: append the arg-max token max_text_len-1 times and return the logits of every position.  The
        reference re-runs decode_text on the growing prefix and once more on the full sequence; with the decoder's
        self-attention K/V cached the per-step logits ARE the logits of that final pass (causal mask), so each token is
        decoded once (Engine.decode_step).  vocab_mask ([V], 0 / -10000) is added before every arg-max and to the
        returned logits (gpv.py:186-187, 193-194)."""
        eng = self.engine
        B, L = s["B"], self.cfg.max_text_len
        st = eng.decode_begin(s["memory"], B, s["Tm"], L)
        tok = torch.full((B,), self.word_to_idx["__cls__"], dtype=torch.int64, device=eng.dev)
        vm = vocab_mask.to(eng.dev).float().contiguous() if vocab_mask is not None else None
        out = torch.empty((B, L, eng.V), device=eng.dev, dtype=torch.float32)
        for t in range(L):
            # masked logits of position t written in place and the first arg-max of them: one launch (gpvb200_argmax)
            tok = k.argmax(eng.decode_step(st, tok), eng.V, vocab_mask=vm, out=out[:, t])
        return out.unsqueeze(0)

    def _graphed(self, kind, images, qids, vocab_mask, beam_size):
        """Whole-call CUDA graph of the inference path (encode + every KV-cached decode step + the on-device arg-max /
        beam bookkeeping): ~1600 launches of 3-10 us each are otherwise issued one by one from Python.  One graph per
        (kind, image shape / dtype, query length, max_text_len, beam size, mask presence); inputs are copied into the
        graph's static buffers, outputs are returned as copies."""
        eng = self.engine
        dev = eng.dev
        key = (kind, tuple(images.shape), images.dtype, tuple(qids.shape), self.cfg.max_text_len, beam_size, vocab_mask is not None)
        g = self._inf_graphs.get(key)
        if g is None:
            st_img, st_q = images.clone(), qids.clone()
            st_vm = vocab_mask.to(dev).float().clone() if vocab_mask is not None else None

            def run():
                s = eng.encode(st_img, st_q, save=False)
                out = self._outputs(s)
                if kind == "greedy":
                    out["answer_logits"] = self._greedy(s, st_vm)
                else:
                    out["beam_token_ids"], out["beam_scores"] = self._beam(s, beam_size)
                return out

            eng.refresh()
            with torch.no_grad():
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    run()                                     # first-call work outside the capture
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    outs = run()
            g = self._inf_graphs[key] = (graph, st_img, st_q, st_vm, outs)
        graph, st_img, st_q, st_vm, outs = g
        st_img.copy_(images, non_blocking=True)
        st_q.copy_(qids, non_blocking=True)
        if st_vm is not None:
            st_vm.copy_(vocab_mask.to(dev).float(), non_blocking=True)
        eng.refresh()
        graph.replay()
        return {k_: v.clone() for k_, v in outs.items()}

    def forward_beam_search(self, images, queries, beam_size=1):
        eng = self.engine
        (images, mask), qids = self._images(images), self._queries(queries)
        eng.train_mode = False
        with torch.no_grad():
            if self.inference_graphs and mask is None:
                outputs = self._graphed("beam", images, qids, None, beam_size)
                seqs, logp = outputs.pop("beam_token_ids"), outputs.pop("beam_scores")
            else:
                s = eng.encode(images, qids, save=False, mask=mask)
                outputs = self._outputs(s)
                seqs, logp = self._beam(s, beam_size)
        seqs_h, p_h = seqs.cpu().tolist(), logp.exp().cpu().tolist()
        stop = {self.word_to_idx["__stop__"], self.word_to_idx["__pad__"]}
        answers = []
        for b in range(len(seqs_h)):
            answers.append([])
            for kk in range(beam_size):
                words = []
                for t in seqs_h[b][kk]:
                    if t in stop:
                        break
                    words.append(self.vocab[t])
                answers[b].append(words)
        outputs["answers"] = answers
        outputs["answer_probs"] = p_h
        outputs["beam_token_ids"] = seqs
        return outputs

    def _beam(self, s, K):
        """gpv.py:256-328 with the beams folded into the batch (row b*K + k), one KV-cached decode step per token and
        the bookkeeping on the device.  Candidate order (k1 major, k2 minor), first-occurrence tie break (stable sort),
        beam 0 only at t = 0, __stop__ never ends a beam -- as the reference."""
        eng = self.engine
        B, Tm, L = s["B"], s["Tm"], self.cfg.max_text_len
        st = eng.decode_begin(s["memory"], B, Tm, L, rep=K)
        dev = eng.dev
        ids = [torch.full((B, K, L), self.word_to_idx["__cls__"], dtype=torch.int64, device=dev) for _ in range(2)]
        score = [torch.zeros((B, K), device=dev), torch.zeros((B, K), device=dev)]
        parent = torch.empty((B * K,), dtype=torch.int64, device=dev)
        work = torch.empty((B * K * K,), dtype=torch.int64, device=dev)       # candidates between the two launches of a beam update
        tok = [ids[0][:, :, 0].reshape(-1).contiguous(), torch.empty((B * K,), dtype=torch.int64, device=dev)]
        for t in range(L - 1):
            lg = eng.decode_step(st, tok[t & 1])
            # log-softmax, per-hypothesis top-K, candidate merge, sequence / score / parent update: gpvb200_beam_update (two launches)
            k.beam_update(lg, eng.V, t, score[t & 1], ids[t & 1], score[(t + 1) & 1], ids[(t + 1) & 1], parent, tok[(t + 1) & 1], work)
            if t < L - 2:
                eng.decode_reorder(st, parent)
        return ids[(L - 1) & 1][:, :, 1:], score[(L - 1) & 1]

    # ------------------------------------------------------------------------------------------------ answers (host)
    def encode_answers(self, targets):
        """gpv.py:377-430 (generation / classification answer encoding; whitespace + punctuation tokenisation stands in
        for nltk.word_tokenize when nltk is absent)."""
        from .tokenizer import word_tokenize
        answers = [t.get("answer", "") for t in targets]
        w2i = self.word_to_idx
        unk = w2i["__unk__"]
        if self.cfg.answering_type == "classification":
            padded = [["__cls__", a] for a in answers]
            ids = [[w2i.get(tok, unk) for tok in row] for row in padded]
        elif self.cfg.answering_type == "generation":
            padded = []
            for a in answers:
                sent = "__cls__ __stop__" if a == "" else f"__cls__ {a} __stop__"
                padded.append([w.lower() for w in word_tokenize(sent)])
            S = max(len(p) for p in padded)
            ids = []
            for p in padded:
                p.extend(["__pad__"] * (S - len(p)))
                ids.append([w2i.get(tok, unk) for tok in p][: self.cfg.max_text_len])
        else:
            raise NotImplementedError
        return padded, torch.as_tensor(ids, dtype=torch.int64).to(self.vision_token.device)

    def token_ids_to_words(self, token_ids):
        ids = token_ids.tolist() if torch.is_tensor(token_ids) else token_ids
        return [[self.vocab[j] for j in row] for row in ids]

    @property
    def cls_token(self):
        raise NotImplementedError("cls_token (unused by the reference's own call sites) is not exposed")
