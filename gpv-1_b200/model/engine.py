"""Explicit forward/backward execution of the GPV-1 hot path on the sm_100a kernels (no autograd graph inside).

`Engine` owns the derived state of one model replica:
  * packed bf16 copies of every weight (FrozenBN folded into the conv weights, q/k/v projections concatenated,
    3x3 convs in [tap][Cout][Cin] order) refreshed by ONE multi-tensor kernel when a parameter version changes;
  * one flat fp32 gradient arena holding the gradient of every parameter that can receive one (the reference
    needs DDP(find_unused_parameters=True) for the rest: train_distr.py:192-193) -- this is also the buffer the
    data-parallel all-reduce runs on (parallel.py);
  * the activations a training step has to keep for its backward, as bf16 token-major / NHWC tensors.

forward_train()/backward() are hand-scheduled sequences of C-ABI calls: every GEMM/conv is the tcgen05 kernel with
bias / residual / activation / ReLU-mask / GELU' fused in its epilogue, attention and LayerNorm are single
kernels, and the matcher + criterion run on the device, so a step never synchronises with the host.

Reference call stack this replaces: GPV.forward gpv.py:137-207 -> DETR.forward detr_roi_head.py:58-94 ->
Backbone backbone.py:71-79, Transformer transformer.py:46-58 -> BertConnectionLayer vilbert.py:872-900 ->
decode_text gpv.py:449-466 -> GPVCriterion losses.py:155-176 -> SetCriterion set_criterion.py:150-191 ->
HungarianMatcher matcher.py:32-77, and autograd's backward of all of it.
"""
import contextlib
import math
import os
import re
import zlib

import torch

from .. import kernels as k
from ..convops import conv_dgrad
from .spec import N_STAGES, grad_stage, never_gets_grad, resnet_blocks

BF16 = torch.bfloat16
F32 = torch.float32
RELU, GELU, SIGMOID = k.ACT_RELU, k.ACT_GELU, k.ACT_SIGMOID
MASK_RELU, GRAD_GELU = k.AUX_RELU_MASK, k.AUX_GELU_GRAD
TASK_LOSS = {"CocoCaptioning": "loss_caption", "CocoVqa": "loss_vqa", "CocoClassification": "loss_cls"}
BB = "detr.backbone.0.body"
_QKV_GROUP = re.compile(r"(co_att_transformer\.\d+)\.qkv([12])\.weight")


def sine_position_table(H, W, device, num_pos_feats=128, temperature=10000.0):
    """position_encoding.py:28-48 for an unpadded H x W map (mask all False): constant, built once per size."""
    y = torch.arange(1, H + 1, dtype=F32, device=device)[:, None].expand(H, W)
    x = torch.arange(1, W + 1, dtype=F32, device=device)[None, :].expand(H, W)
    y = y / (H + 1e-6) * (2 * math.pi)
    x = x / (W + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=F32, device=device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(H * W, 2 * num_pos_feats)


def sine_position_masked(mask, num_pos_feats=128, temperature=10000.0):
    """position_encoding.py:28-48 for a padded batch: mask [B,H,W] bool (True = padding) -> [B*H*W, 2*num_pos_feats] fp32.
    Tiny torch ops (cumsum over a 15x20 map per image); only used when a batch mixes image sizes."""
    nm = ~mask
    y = nm.cumsum(1, dtype=F32)
    x = nm.cumsum(2, dtype=F32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=F32, device=mask.device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).reshape(-1, 2 * num_pos_feats)


class _Lane:
    """Context of Engine._aside: work issued inside runs on the lane's stream, and its atomically accumulating GEMM launches (weight
    gradients) get the lane's CTA cap (kernels.LANE_CTA_CAP -> gpvb200_gemm_desc.max_ctas)."""

    def __init__(self, stream, cap):
        self.ctx, self.cap = torch.cuda.stream(stream), cap

    def __enter__(self):
        self.prev, k.LANE_CTA_CAP = k.LANE_CTA_CAP, self.cap
        return self.ctx.__enter__()

    def __exit__(self, *a):
        k.LANE_CTA_CAP = self.prev
        return self.ctx.__exit__(*a)


class Engine:
    def __init__(self, tensors, specs, cfg, device):
        """tensors: name -> fp32 CUDA tensor for every state_dict entry (nn.Parameters and buffers of the module that
        owns this engine); specs: model.spec.gpv_specs(); cfg: the `model:` config block."""
        self.P = tensors
        self.cfg = cfg
        self.dev = device
        self.d = cfg.detr.hidden_dim
        self.D = cfg.hidden_dim
        self.Q = cfg.detr.num_queries
        self.n_enc, self.n_dec = cfg.detr.num_encoder_layers, cfg.detr.num_decoder_layers
        self.n_co, self.n_txt = cfg.co_att.num_layers, cfg.text_decoder.num_layers
        self.h_detr, self.h_co, self.h_txt = cfg.detr.nheads, cfg.co_att.bi_num_attention_heads, cfg.text_decoder.nheads
        self.V = tensors["answer_head.vocab_embed"].shape[0]
        self.Vp = (self.V + 7) // 8 * 8
        self.blocks = list(resnet_blocks())
        self.unit_upstream_grad = True
        lw = {}
        for _, lc in cfg.losses.items():
            lw.update(lc.loss_wts)
        self.loss_wts = {kk: float(v) for kk, v in lw.items()}
        loc = cfg.losses.Localization
        self.cost_w = (float(loc.cost_wts.ce), float(loc.cost_wts.bbox), float(loc.cost_wts.giou))
        self.eos_coef = float(loc.eos_coef)
        lwv = [1.0, self.loss_wts["loss_ce"], self.loss_wts["loss_bbox"], self.loss_wts["loss_giou"]]
        self._wts_loc = torch.tensor(lwv, dtype=F32, device=device)
        self._wts_noloc = torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=F32, device=device)
        self._setup_weights()
        self._setup_grads(specs)
        self._versions = None
        self._pos_cache = {}
        self.saved = None
        # train-mode dropout (every nn.Dropout site of the reference uses cfg.detr.dropout = 0.1: gpv.yaml:50,64,75-82;
        # HF BERT 0.1).  train_mode is set by the owning module from nn.Module.training before each call.
        self.train_mode = False
        co = cfg.co_att
        self.p_drop = {"detr": float(cfg.detr.dropout), "txt": float(cfg.text_decoder.dropout), "bert": 0.1,
                       "co_att_l": float(co.v_attention_probs_dropout_prob), "co_att_v": float(co.attention_probs_dropout_prob),
                       "co_hid_l": float(co.v_hidden_dropout_prob), "co_hid_v": float(co.hidden_dropout_prob)}
        # training-step counter read by the dropout kernels; the rank sits in the high bits so that data-parallel replicas draw
        # different masks (torch's per-process Philox streams differ too); train.save_checkpoint / load_checkpoint carry it
        import torch.distributed as _dist
        rank = _dist.get_rank() if _dist.is_available() and _dist.is_initialized() else 0
        self.drop_seed = torch.full((1,), rank << 40, dtype=torch.int64, device=device)
        self.frozen = set()                                   # names with requires_grad = False: no weight / bias gradient kernels
        self.last_stage = N_STAGES - 1                        # last gradient stage backward() reaches (lower when the tail is frozen)
        self.concurrent = True                                # run independent branches on side streams (lanes)
        self.fused_layers = os.environ.get("GPVB200_FUSED", "1") != "0"   # row-tile-resident sub-layer kernels (layer_umma.cu)
        # fused FFN data gradients (gpvb200_mlp_block_bwd): parity-green, but inside the step its 75 row tiles lose 0.08 ms against the
        # two 148-SM GEMMs they replace (profiles/r2z_mlp_block_bwd.txt), so it is opt-in
        self.fused_bwd = self.fused_layers and os.environ.get("GPVB200_MLP_BWD", "0") == "1"
        self.use_attn_block = self.fused_layers and os.environ.get("GPVB200_ATTN_BLOCK", "1") != "0"   # tcgen05 attention + out-proj + LN
        # lane scheduling (measured in profiles/r2q_lanes.txt: 16.97 -> 16.51 ms per step): the weight-gradient lanes may lag behind the
        # data-gradient chain until the end of a gradient stage instead of being joined after every layer / bottleneck
        # (GPVB200_LAZY_JOIN=0 restores the per-layer join), and the lane-0 work (weight and bias gradients) is dealt round-robin to
        # GPVB200_WGRAD_LANES streams (default 3) so that independent weight gradients overlap each other
        # weight-gradient GEMMs issued on a lane use at most this many CTAs (half the SMs: the data-gradient chain keeps the rest;
        # GPVB200_WGRAD_CTAS overrides, 0 = uncapped; scan in profiles/r2t_wgrad_cta_cap.txt)
        self.lane_cta_cap = int(os.environ.get("GPVB200_WGRAD_CTAS", str(torch.cuda.get_device_properties(device).multi_processor_count // 2)))
        self.lazy_join = os.environ.get("GPVB200_LAZY_JOIN", "1") == "1"
        self.n_wlanes = max(1, int(os.environ.get("GPVB200_WGRAD_LANES", "3")))
        self._rr = 0
        self._lanes, self._dirty, self._keep = {}, set(), []

    # ================================================================================================ weights
    def _setup_weights(self):
        P, dev = self.P, self.dev
        items, W, self.Bcat_src, self.Bcat = [], {}, {}, {}

        def lin(name, key=None):
            t = P[name]
            N, K = t.shape[0], t.numel() // t.shape[0]
            dst = torch.empty((N, K), device=dev, dtype=BF16)
            items.append((t, dst, None, N, K, 1, 0))
            W[key or name] = dst

        def linT(name):
            """bf16 W^T [K, N] of an nn.Linear weight [N, K] (the pack kernel's tap index walks the input features)."""
            t = P[name]
            N, K = t.shape
            dst = torch.empty((K, N), device=dev, dtype=BF16)
            items.append((t, dst, None, N, 1, K, 0))
            W[name + ".T"] = dst

        def cat(key, names):
            K = P[names[0] + ".weight"].shape[1]
            Ns = [P[n + ".weight"].shape[0] for n in names]
            dst = torch.empty((sum(Ns), K), device=dev, dtype=BF16)
            off = 0
            for n, N in zip(names, Ns):
                items.append((P[n + ".weight"], dst[off:off + N], None, N, K, 1, 0))
                off += N
            W[key] = dst
            self.Bcat_src[key] = [P[n + ".bias"] for n in names]
            self.Bcat[key] = torch.empty(sum(Ns), device=dev, dtype=F32)

        # ---- backbone: FrozenBN folded (scale into the packed weight rows, bias into the epilogue)
        bns = [f"{BB}.bn1"]
        for li, bi, inp, planes, s, ds in self.blocks:
            p = f"{BB}.layer{li}.{bi}"
            bns += [p + ".bn1", p + ".bn2", p + ".bn3"] + ([p + ".downsample.1"] if ds else [])
        tot = sum(P[b + ".weight"].numel() for b in bns)
        self._bn_flat = torch.empty(2 * tot, device=dev, dtype=F32)
        self.bn_scale, self.bn_bias, off = {}, {}, 0
        for b in bns:
            n = P[b + ".weight"].numel()
            self.bn_scale[b] = self._bn_flat[off:off + n]
            self.bn_bias[b] = self._bn_flat[tot + off:tot + off + n]
            off += n
        self.bns = bns

        def conv(name, bn):
            t = P[name]
            O, I, kh, kw = t.shape
            dst = torch.empty((kh * kw, O, I), device=dev, dtype=BF16)
            items.append((t, dst, self.bn_scale[bn], O, I, kh * kw, 0))
            W[name] = dst

        stem = torch.empty((64, 152), device=dev, dtype=BF16)
        items.append((P[f"{BB}.conv1.weight"], stem, self.bn_scale[f"{BB}.bn1"], 64, 3, 49, 1))
        W[f"{BB}.conv1.weight"] = stem
        for li, bi, inp, planes, s, ds in self.blocks:
            p = f"{BB}.layer{li}.{bi}"
            conv(p + ".conv1.weight", p + ".bn1")
            conv(p + ".conv2.weight", p + ".bn2")
            conv(p + ".conv3.weight", p + ".bn3")
            if ds:
                conv(p + ".downsample.0.weight", p + ".downsample.1")
        # ---- DETR
        lin("detr.input_proj.weight")
        lin("detr.query_embed.weight")
        for i in range(self.n_enc):
            p = f"detr.transformer.encoder.layers.{i}"
            for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "linear1.weight", "linear2.weight"):
                lin(f"{p}.{n}")
            linT(f"{p}.linear1.weight")       # K-major operands of the fused FFN backward (gpvb200_mlp_block_bwd)
            linT(f"{p}.linear2.weight")
        for i in range(self.n_dec):
            p = f"detr.transformer.decoder.layers.{i}"
            for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "multihead_attn.in_proj_weight",
                      "multihead_attn.out_proj.weight", "linear1.weight", "linear2.weight"):
                lin(f"{p}.{n}")
        for n in ("detr.class_embed", "detr.bbox_embed.layers.0", "detr.bbox_embed.layers.1", "detr.bbox_embed.layers.2",
                  "detr_joiner", "bert_joiner", "relevance_predictor", "answer_head.classifier_transform",
                  "answer_input_embedings.transform"):
            lin(n + ".weight")
        lin("answer_head.vocab_embed")
        # ---- BERT (forward only)
        for i in range(12):
            p = f"bert.model.encoder.layer.{i}"
            cat(p + ".qkv", [p + ".attention.self.query", p + ".attention.self.key", p + ".attention.self.value"])
            for n in ("attention.output.dense", "intermediate.dense", "output.dense"):
                lin(f"{p}.{n}.weight")
        # ---- co-attention
        for i in range(self.n_co):
            p = f"co_att_transformer.{i}"
            cat(p + ".qkv1", [f"{p}.biattention.{n}1" for n in ("query", "key", "value")])
            cat(p + ".qkv2", [f"{p}.biattention.{n}2" for n in ("query", "key", "value")])
            for n in ("biOutput.dense1", "biOutput.dense2", "v_intermediate.dense", "v_output.dense", "t_intermediate.dense",
                      "t_output.dense"):
                lin(f"{p}.{n}.weight")
        # ---- text decoder
        for i in range(self.n_txt):
            p = f"text_decoder.layers.{i}"
            for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "multihead_attn.in_proj_weight",
                      "multihead_attn.out_proj.weight", "linear1.weight", "linear2.weight"):
                lin(f"{p}.{n}")
        self.W = W
        self._plan = k.PackPlan(items, dev)

    def mark_dirty(self):
        """The trainable parameters were updated behind torch's version counters (optim.ClipAdamW writes them through raw
        pointers): re-derive the packed bf16 copies at the next forward."""
        self._versions = None

    @torch.no_grad()
    def refresh(self, force=False):
        """Re-derive the packed bf16 weights when any parameter changed (optimizer step, load_state_dict).  The frozen
        pieces (FrozenBatchNorm buffers -> folded scale / bias, the stem convolution) are tracked separately so that an
        optimizer step costs one multi-tensor pack launch, not 53 BN folds."""
        P = self.P
        ver = sum(t._version for t in P.values())
        if not force and ver == self._versions:
            return
        frozen = [P[b + sfx] for b in self.bns for sfx in (".weight", ".bias", ".running_mean", ".running_var")] + [P[f"{BB}.conv1.weight"]]
        fver = sum(t._version for t in frozen)
        if force or fver != getattr(self, "_frozen_versions", None):
            for b in self.bns:
                k.bn_fold(P[b + ".weight"], P[b + ".bias"], P[b + ".running_mean"], P[b + ".running_var"], self.bn_scale[b],
                          self.bn_bias[b])
            self.W["stem.s2d"] = k.stem_weight_s2d(P[f"{BB}.conv1.weight"] * self.bn_scale[f"{BB}.bn1"].view(-1, 1, 1, 1))
            self._frozen_versions = fver
        self._plan.run()
        for key, srcs in self.Bcat_src.items():
            torch.cat(srcs, out=self.Bcat[key])
        self._versions = sum(t._version for t in P.values())

    # ================================================================================================ gradients
    def _setup_grads(self, specs):
        live = [s for s in specs if s.kind == "param" and not never_gets_grad(s.name)]
        live.sort(key=lambda s: grad_stage(s.name))          # stable: spec order within a stage
        byname = {s.name: s for s in live}
        groups = []
        for i in range(self.n_co):
            p = f"co_att_transformer.{i}.biattention"
            for sfx in ("1", "2"):
                groups.append([f"{p}.{n}{sfx}.weight" for n in ("query", "key", "value")])
                groups.append([f"{p}.{n}{sfx}.bias" for n in ("query", "key", "value")])
        first = {g[0]: g for g in groups}
        member = {n for g in groups for n in g}
        order = []
        for s in live:
            if s.name in first:
                order += first[s.name]
            elif s.name not in member:
                order.append(s.name)
        total = 0
        offs = {}
        for n in order:
            offs[n] = total
            total += (math.prod(byname[n].shape) + 7) // 8 * 8 if n not in member else math.prod(byname[n].shape)
        total = (total + 7) // 8 * 8
        self.stage_end = [0] * N_STAGES                       # arena offset (elements) where each stage's gradients end
        for n in order:
            st = grad_stage(n)
            self.stage_end[st] = max(self.stage_end[st], (offs[n] + math.prod(byname[n].shape) + 7) // 8 * 8)
        for st in range(1, N_STAGES):
            self.stage_end[st] = max(self.stage_end[st], self.stage_end[st - 1])
        self.stage_end[-1] = total
        self.on_stage_done = None                             # parallel.py: callable(stage) fired as backward finishes a stage
        self.on_backward_end = None                           # parallel.py: joins the all-reduce stream (inside a graph capture too)
        self.stage_joins = None                               # parallel.py: callable(stage) -> does a bucket end (and get reduced) after this stage?
        self.grad_arena = torch.zeros(total, device=self.dev, dtype=F32)
        self.G = {n: self.grad_arena[offs[n]:offs[n] + math.prod(byname[n].shape)].view(byname[n].shape) for n in order}
        self.live_names = order
        for i in range(self.n_co):
            p = f"co_att_transformer.{i}"
            for sfx in ("1", "2"):
                names = [f"{p}.biattention.{n}{sfx}" for n in ("query", "key", "value")]
                o = offs[names[0] + ".weight"]
                Dm = byname[names[0] + ".weight"].shape
                self.G[f"{p}.qkv{sfx}.weight"] = self.grad_arena[o:o + 3 * Dm[0] * Dm[1]].view(3 * Dm[0], Dm[1])
                o = offs[names[0] + ".bias"]
                self.G[f"{p}.qkv{sfx}.bias"] = self.grad_arena[o:o + 3 * Dm[0]]
        # packed [taps,O,I] accumulators for the 3x3 convs (unpacked into the Conv2d layout at the end of backward)
        n3 = 0
        self._g3 = []
        for li, bi, inp, planes, s, ds in self.blocks:
            if li >= 2:
                self._g3.append((f"{BB}.layer{li}.{bi}.conv2.weight", planes, n3))
                n3 += 9 * planes * planes
        self.grad_pack = torch.zeros(n3, device=self.dev, dtype=F32)
        self.Gp = {n: self.grad_pack[o:o + 9 * c * c].view(9, c, c) for n, c, o in self._g3}

    # ================================================================================================ small helpers
    def _done(self, stage):
        # the lanes are joined where a gradient bucket is handed to the all-reduce (a hook is attached) and where backward ends;
        # without a hook (one GPU) intermediate stage boundaries do not make the data-gradient chain wait for the lanes
        hook_joins = self.on_stage_done is not None and (self.stage_joins is None or self.stage_joins(stage))
        if hook_joins or stage >= self.last_stage or not self.lazy_join:
            self._join()
        if self.on_stage_done is not None:
            self.on_stage_done(stage)

    # ---- dropout sites
    def _pfam(self, name):
        if name.startswith("detr."):
            return self.p_drop["detr"]
        if name.startswith("text_decoder."):
            return self.p_drop["txt"]
        if name.startswith("bert."):
            return self.p_drop["bert"]
        return 0.0

    def _drop(self, site, p=None):
        """k.Drop for the dropout site named `site` (None in eval mode or when p = 0).  Forward and backward of a site
        call this with the same name, so both regenerate the same mask from (step counter, crc32(name))."""
        if not self.train_mode:
            return None
        p = self._pfam(site) if p is None else p
        if p <= 0.0:
            return None
        return k.Drop(self.drop_seed, zlib.crc32(site.encode()), p)

    def _trains(self, name):
        """False when the parameter (or, for the packed co-attention q/k/v groups, its first member) is frozen."""
        if not self.frozen:
            return True
        if name in self.frozen:
            return False
        m = _QKV_GROUP.match(name)
        if m:
            return f"{m.group(1)}.biattention.query{m.group(2)}.weight" not in self.frozen
        return True

    def _backward_depth(self):
        """How far down the DETR sub-graph backward() has to go, from which parameters still train (freeze_detr_params,
        train_distr.py:136-140, freezes whatever was initialised from the DETR checkpoint): 4 = backbone, 3 = encoder /
        input_proj, 2 = decoder layers / query_embed, 1 = box / class heads and the decoder's final norm, 0 = nothing
        under `detr.` trains, so no data gradient enters it at all (only detr_joiner's weight gradient is left)."""
        if not self.frozen:
            return 4
        depth = 0
        for n in self.G:
            if n in self.frozen or not n.startswith("detr."):
                continue
            if n.startswith("detr.backbone."):
                return 4
            if n.startswith(("detr.transformer.encoder.", "detr.input_proj.")):
                depth = max(depth, 3)
            elif n.startswith("detr.transformer.decoder.layers.") or n == "detr.query_embed.weight":
                depth = max(depth, 2)
            else:
                depth = max(depth, 1)
        return depth

    def _backward_end(self):
        if self.on_backward_end is not None:
            self.on_backward_end()
        return self.G

    def _ln_bwd(self, dy, x, st, gamma, dgamma, dbeta, drop):
        """(dx, dx_masked): LayerNorm backward of y = LN(res + dropout(f)); dx_masked is the gradient of f's output."""
        if drop is None:
            dx = k.layernorm_bwd(dy, x, st, gamma, dgamma, dbeta)
            return dx, dx
        return k.layernorm_bwd(dy, x, st, gamma, dgamma, dbeta, drop=drop)

    # ---- concurrent branches.  Most GEMMs of the transformer stacks cover 15-75 of the 148 SMs, so independent
    # branches (weight/bias gradients beside the data-gradient chain, BERT beside the backbone, K/V projections of a
    # fixed memory beside the decoder's self-attention) are enqueued on side streams ("lanes"); inside a captured
    # step they become parallel branches of the CUDA graph.  Rule that keeps the caching allocator safe: every tensor
    # allocated on the issuing stream and read on a lane is passed to _aside() and stays referenced until _join().
    def _aside(self, *keep, lane=0):
        if not self.concurrent or self.dev.type != "cuda":
            return contextlib.nullcontext()
        if lane == 0 and self.n_wlanes > 1:
            lane = 100 + self._rr % self.n_wlanes
            self._rr += 1
        st = self._lanes.get(lane)
        if st is None:
            st = self._lanes[lane] = torch.cuda.Stream(device=self.dev)
        st.wait_stream(torch.cuda.current_stream())
        self._keep.extend(t for t in keep if t is not None)
        self._dirty.add(lane)
        return _Lane(st, self.lane_cta_cap)

    def _aside_after_lanes(self, lane):
        """A lane whose work also follows everything issued on the other lanes so far (not only the issuing stream)."""
        if not (self.concurrent and self.lazy_join and self.dev.type == "cuda"):
            self._join()
            return contextlib.nullcontext()
        ctx = self._aside(lane=lane)
        st = self._lanes[lane]
        for ln in list(self._dirty):
            if ln != lane:
                st.wait_stream(self._lanes[ln])
        return ctx

    def _join_layer(self, *keep):
        """End of a layer / bottleneck of the backward pass: join the lanes, or (lazy) only keep what they still read alive."""
        if self.lazy_join and self.concurrent and self.dev.type == "cuda":
            self._keep.extend(t for t in keep if t is not None)
        else:
            self._join()

    def _join(self, lane=None):
        lanes = list(self._dirty) if lane is None else ([lane] if lane in self._dirty else [])
        cur = torch.cuda.current_stream() if lanes else None
        for ln in lanes:
            cur.wait_stream(self._lanes[ln])
            self._dirty.discard(ln)
        if not self._dirty:
            self._keep.clear()

    def _pos(self, H, W):
        key = (H, W)
        if key not in self._pos_cache:
            self._pos_cache[key] = k.cast_bf16(sine_position_table(H, W, self.dev))
        return self._pos_cache[key]

    def _lin_bwd(self, name, x, dy, *, need_dx=True, aux=None, aux_mode=k.AUX_NONE, residual=None, wkey=None, gkey=None,
                 bias=True, alpha=1.0):
        """dW += dy^T x, db += colsum(dy), returns dx = (dy W + residual) (*) mask."""
        g = gkey or name
        if self._trains(g + ".weight"):
            with self._aside(dy, x):
                k.linear_wgrad(dy, x, self.G[g + ".weight"].view(dy.shape[1], -1))
                if bias:
                    k.colsum(dy, self.G[g + ".bias"])
        if need_dx:
            return k.linear_dgrad(dy, self.W[wkey or (name + ".weight")], aux=aux, aux_mode=aux_mode, residual=residual, alpha=alpha)
        return None

    # ================================================================================================ backbone
    def _backbone_fwd(self, images, save):
        W, bb = self.W, self.bn_bias
        B = images.shape[0]
        xv, Ho, Wo = k.stem_s2d(images)                       # 7x7/s2 stem as a 4-tap K=64 implicit GEMM over the s2d map
        x = k.conv(xv, W["stem.s2d"], ksize=7, taps=k.STEM_TAPS, Ho=Ho, Wo=Wo, N=64, K=64, bias=bb[f"{BB}.bn1"], act=RELU)
        del xv
        x = k.maxpool3x3s2(x)
        acts = []
        for blk in self.blocks:
            y, saved = self._bottleneck_fwd(blk, x)
            if save and blk[0] >= 2:
                acts.append(saved)
            x = y
        return x, acts

    def _bottleneck_fwd(self, blk, x):
        """torchvision Bottleneck (v1.5) with FrozenBN folded: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1 x))))))) + idn)."""
        W, bb = self.W, self.bn_bias
        li, bi, inp, planes, s, ds = blk
        p = f"{BB}.layer{li}.{bi}"
        n, H, Wd, _ = x.shape
        h1 = k.linear(x.view(-1, inp), W[p + ".conv1.weight"][0], bb[p + ".bn1"], act=RELU).view(n, H, Wd, planes)
        h2 = k.conv(h1, W[p + ".conv2.weight"], ksize=3, stride=s, bias=bb[p + ".bn2"], act=RELU)
        Ho2, Wo2 = h2.shape[1], h2.shape[2]
        if ds:
            if s == 1:
                idn = k.linear(x.view(-1, inp), W[p + ".downsample.0.weight"][0], bb[p + ".downsample.1"]).view(n, H, Wd, planes * 4)
            else:
                idn = k.conv(x, W[p + ".downsample.0.weight"], ksize=1, stride=s, bias=bb[p + ".downsample.1"])
        else:
            idn = x
        y = k.linear(h2.view(-1, planes), W[p + ".conv3.weight"][0], bb[p + ".bn3"], residual=idn.view(-1, planes * 4),
                     act=RELU).view(n, Ho2, Wo2, planes * 4)
        return y, (x, h1, h2, y)

    def _bottleneck_bwd(self, blk, dpre, saved, need_dx):
        """Backward of one torchvision Bottleneck with FrozenBN.  dpre: gradient w.r.t. the block's pre-ReLU output
        (already masked by relu'(y)).  Returns the masked gradient w.r.t. the previous block's pre-ReLU output."""
        W, G, sc = self.W, self.G, self.bn_scale
        li, bi, inp, planes, s, ds = blk
        p = f"{BB}.layer{li}.{bi}"
        x, h1, h2, y = saved
        n, H, Wd, _ = x.shape
        dpre2 = dpre.view(-1, planes * 4)
        h2f, h1f, xf = h2.view(-1, planes), h1.view(-1, planes), x.view(-1, inp)
        # conv3 (1x1)
        dh2 = k.linear_dgrad(dpre2, W[p + ".conv3.weight"][0], aux=h2f, aux_mode=MASK_RELU)
        with self._aside(dpre):
            if self._trains(p + ".conv3.weight"):
                k.linear_wgrad(dpre2, h2f, G[p + ".conv3.weight"].view(planes * 4, planes), rowscale=sc[p + ".bn3"])
            if ds and self._trains(p + ".downsample.0.weight"):
                gds = G[p + ".downsample.0.weight"].view(1, planes * 4, inp)
                if s == 1:
                    k.linear_wgrad(dpre2, xf, gds[0], rowscale=sc[p + ".downsample.1"])
                else:
                    k.conv_wgrad(dpre, x, gds, ksize=1, stride=s, rowscale=sc[p + ".downsample.1"])
        # conv2 (3x3, stride s)
        dh2 = dh2.view(h2.shape)
        dh1 = conv_dgrad(dh2, W[p + ".conv2.weight"], ksize=3, stride=s, in_hw=(H, Wd), aux=h1, aux_mode=MASK_RELU)
        if self._trains(p + ".conv2.weight"):
            with self._aside(dh2):
                k.conv_wgrad(dh2, h1, self.Gp[p + ".conv2.weight"], ksize=3, stride=s, rowscale=sc[p + ".bn2"])
        # conv1 (1x1) + identity / downsample branch: dx = (dh1 W1 + didn) * relu'(x)  -> already the masked gradient of the previous block
        if self._trains(p + ".conv1.weight"):
            with self._aside(dh1):
                k.linear_wgrad(dh1.view(-1, planes), xf, G[p + ".conv1.weight"].view(planes, inp), rowscale=sc[p + ".bn1"])
        dx = None
        if need_dx and ds and s == 2:
            # stride-2 1x1 down-sample: its data gradient touches only the even positions of dx.  dx = (dh1 W1) * relu'(x) first, then
            # ONE implicit-GEMM launch adds dpre Wds^T at those positions in place (residual = out; the mask is idempotent), instead
            # of materialising the mostly-zero branch gradient (three strided zero-fills + one GEMM + a full-size residual read)
            dx = k.linear_dgrad(dh1.view(-1, planes), W[p + ".conv1.weight"][0], aux=xf, aux_mode=MASK_RELU).view(x.shape)
            conv_dgrad(dpre, W[p + ".downsample.0.weight"], ksize=1, stride=2, in_hw=(H, Wd), aux=x, aux_mode=MASK_RELU, residual=dx,
                       out=dx, only_parity=(0, 0))
        elif need_dx:
            didn = conv_dgrad(dpre, W[p + ".downsample.0.weight"], ksize=1, stride=s, in_hw=(H, Wd)) if ds else dpre
            dx = k.linear_dgrad(dh1.view(-1, planes), W[p + ".conv1.weight"][0], residual=didn.view(-1, inp), aux=xf,
                                aux_mode=MASK_RELU).view(x.shape)
        self._join_layer(*saved)            # the saved activations of this block are released by the caller
        return dx

    def _backbone_bwd(self, dpre, acts):
        """dpre: gradient w.r.t. the pre-ReLU output of the last block (already masked), NHWC bf16."""
        trainable = [b for b in self.blocks if b[0] >= 2]
        for idx in range(len(trainable) - 1, -1, -1):
            blk = trainable[idx]
            li, bi = blk[0], blk[1]
            saved, acts[idx] = acts[idx], None
            dpre = self._bottleneck_bwd(blk, dpre, saved, need_dx=idx > 0)
            if bi == 0:                                        # first block of layer `li` was the last one to finish
                # the packed 3x3 weight gradients are written on the lanes: they are unpacked on a lane that waits for those lanes, so
                # the data-gradient chain itself never waits (_done() joins where a bucket is reduced or backward ends)
                with self._aside_after_lanes(5):
                    for name, c, _ in self._g3:
                        if f".layer{li}." in name:
                            k.unpack_conv_grad(self.Gp[name], self.G[name])
                self._done({4: 4, 3: 5, 2: 6}[li])

    # ================================================================================================ attention blocks
    def _self_attn_fwd(self, p, x, pos, P_rows, B, S, H, *, causal=False, eps=1e-5, norm="norm1", attn="self_attn", kmask=None):
        """y = LN(x + out_proj(MHA(q = k = x + pos, v = x))).  x [B*S, D]; pos [P_rows, D] bf16 or None."""
        W, Pm = self.W, self.P
        D = x.shape[1]
        wi, bi_ = W[f"{p}.{attn}.in_proj_weight"], Pm[f"{p}.{attn}.in_proj_bias"]
        qkv = torch.empty((x.shape[0], 3 * D), device=self.dev, dtype=BF16)
        if pos is not None:
            qk_in = k.add_rowbcast(x, pos)
            k.linear(qk_in, wi[:2 * D], bi_[:2 * D], out=qkv[:, :2 * D])
            k.linear(x, wi[2 * D:], bi_[2 * D:], out=qkv[:, 2 * D:])
        else:
            qk_in = x
            k.linear(x, wi, bi_, out=qkv)
        dh = D // H
        if self.use_attn_block and D == 256 and H == 8 and not causal and S <= 304 and B * ((S + 127) // 128) >= 64:
            # DETR encoder self-attention: tcgen05 attention core + out-proj + dropout + residual + LayerNorm in ONE launch
            # (gpvb200_attn_block_fwd); with fewer than ~64 row tiles (the decoder's 100 queries per image) its 8-heads-per-CTA
            # walk leaves most SMs idle and the (batch, head)-parallel mma.sync kernel below is faster
            y, o, lse, pre, st = k.attn_block_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B=B, Sq=S, Sk=S, scale=dh ** -0.5,
                                                  key_mask=kmask, wo=W[f"{p}.{attn}.out_proj.weight"], bo=Pm[f"{p}.{attn}.out_proj.bias"],
                                                  x=x, gamma=Pm[f"{p}.{norm}.weight"], beta=Pm[f"{p}.{norm}.bias"], eps=eps,
                                                  drop_p=self._drop(f"{p}.{attn}.probs"), drop_o=self._drop(f"{p}.{norm}.in"))
            return y, (x, qk_in, qkv, o, lse, pre, st)
        o, lse = k.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B=B, H=H, Sq=S, Sk=S, dh=dh, scale=dh ** -0.5,
                                 causal=causal, key_mask=kmask, drop=self._drop(f"{p}.{attn}.probs"))
        pre = k.linear(o, W[f"{p}.{attn}.out_proj.weight"], Pm[f"{p}.{attn}.out_proj.bias"], residual=x,
                       drop=self._drop(f"{p}.{norm}.in"), drop_mode=k.DROP_PRE_RESIDUAL)
        y, st = k.layernorm_fwd(pre, Pm[f"{p}.{norm}.weight"], Pm[f"{p}.{norm}.bias"], eps)
        return y, (x, qk_in, qkv, o, lse, pre, st)

    def _self_attn_bwd(self, p, dy, sv, pos_grad, B, S, H, *, causal=False, norm="norm1", attn="self_attn", has_pos=True, kmask=None):
        """Returns dx.  pos_grad: fp32 [S, D] accumulator for a learned position (query_embed) or None."""
        W, Pm, G = self.W, self.P, self.G
        x, qk_in, qkv, o, lse, pre, st = sv
        D = x.shape[1]
        dh = D // H
        dpre, dpre_m = self._ln_bwd(dy, pre, st, Pm[f"{p}.{norm}.weight"], G[f"{p}.{norm}.weight"], G[f"{p}.{norm}.bias"],
                                    self._drop(f"{p}.{norm}.in"))
        do = self._lin_bwd(f"{p}.{attn}.out_proj", o, dpre_m)
        dqkv = torch.empty_like(qkv)
        k.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                        B=B, H=H, Sq=S, Sk=S, dh=dh, scale=dh ** -0.5, causal=causal, key_mask=kmask,
                        drop=self._drop(f"{p}.{attn}.probs"))
        wi = W[f"{p}.{attn}.in_proj_weight"]
        gw, gb = G[f"{p}.{attn}.in_proj_weight"], G[f"{p}.{attn}.in_proj_bias"]
        if self._trains(f"{p}.{attn}.in_proj_weight"):
            with self._aside(dqkv):
                k.colsum(dqkv, gb)
                if not has_pos:
                    k.linear_wgrad(dqkv, x, gw)
                else:
                    k.linear_wgrad(dqkv[:, :2 * D], qk_in, gw[:2 * D])
                    k.linear_wgrad(dqkv[:, 2 * D:], x, gw[2 * D:])
        if not has_pos:
            return k.linear_dgrad(dqkv, wi, residual=dpre)
        if pos_grad is None:
            # fixed (sine) position: d(x + pos)/dx = I, so the q / k path and the v path share ONE data-gradient GEMM over K = 3 D
            return k.linear_dgrad(dqkv, wi, residual=dpre)
        dx = k.linear_dgrad(dqkv[:, 2 * D:], wi[2 * D:], residual=dpre)
        dqk_in = k.linear_dgrad(dqkv[:, :2 * D], wi[:2 * D])
        k.batch_reduce(dqk_in, pos_grad, B, S)
        return k.add(dx, dqk_in, out=dx)

    def _cross_kv(self, p, kmem, vmem):
        """K/V projections of a cross-attention memory: independent of the decoder state, so the callers enqueue them
        for every layer at once on a lane, beside the first layers of the decoder."""
        W, Pm = self.W, self.P
        D = kmem.shape[1]
        wi, bi_ = W[f"{p}.multihead_attn.in_proj_weight"], Pm[f"{p}.multihead_attn.in_proj_bias"]
        kv = torch.empty((kmem.shape[0], 2 * D), device=self.dev, dtype=BF16)
        if kmem is vmem:
            k.linear(kmem, wi[D:], bi_[D:], out=kv)
        else:
            k.linear(kmem, wi[D:2 * D], bi_[D:2 * D], out=kv[:, :D])
            k.linear(vmem, wi[2 * D:], bi_[2 * D:], out=kv[:, D:])
        return kv

    def _cross_attn_fwd(self, p, x, qpos, kmem, vmem, B, Sq, Sk, H, *, eps=1e-5, norm="norm2", kv=None, kmask=None):
        """y = LN(x + out_proj(MHA(q = x + qpos, k = kmem, v = vmem)))  (kmem already carries its position)."""
        W, Pm = self.W, self.P
        D = x.shape[1]
        wi, bi_ = W[f"{p}.multihead_attn.in_proj_weight"], Pm[f"{p}.multihead_attn.in_proj_bias"]
        q_in = k.add_rowbcast(x, qpos) if qpos is not None else x
        q = k.linear(q_in, wi[:D], bi_[:D])
        if kv is None:
            kv = self._cross_kv(p, kmem, vmem)
        dh = D // H
        o, lse = k.attention_fwd(q, kv[:, :D], kv[:, D:], B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=dh ** -0.5, key_mask=kmask,
                                 drop=self._drop(f"{p}.multihead_attn.probs"))
        pre = k.linear(o, W[f"{p}.multihead_attn.out_proj.weight"], Pm[f"{p}.multihead_attn.out_proj.bias"], residual=x,
                       drop=self._drop(f"{p}.{norm}.in"), drop_mode=k.DROP_PRE_RESIDUAL)
        y, st = k.layernorm_fwd(pre, Pm[f"{p}.{norm}.weight"], Pm[f"{p}.{norm}.bias"], eps)
        return y, (x, q_in, q, kv, o, lse, pre, st)

    def _cross_attn_bwd(self, p, dy, sv, kmem, vmem, dmem, pos_grad, B, Sq, Sk, H, *, norm="norm2", kmask=None):
        """Returns (dx, dmem) with dmem = dmem_in + d(kmem) + d(vmem) (chained through the GEMM residual input)."""
        W, Pm, G = self.W, self.P, self.G
        x, q_in, q, kv, o, lse, pre, st = sv
        D = x.shape[1]
        dh = D // H
        a = f"{p}.multihead_attn"
        dpre, dpre_m = self._ln_bwd(dy, pre, st, Pm[f"{p}.{norm}.weight"], G[f"{p}.{norm}.weight"], G[f"{p}.{norm}.bias"],
                                    self._drop(f"{p}.{norm}.in"))
        do = self._lin_bwd(a + ".out_proj", o, dpre_m)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        k.attention_bwd(q, kv[:, :D], kv[:, D:], o, do, lse, dq, dkv[:, :D], dkv[:, D:], B=B, H=H, Sq=Sq, Sk=Sk, dh=dh,
                        scale=dh ** -0.5, key_mask=kmask, drop=self._drop(f"{p}.multihead_attn.probs"))
        wi, gw, gb = W[a + ".in_proj_weight"], G[a + ".in_proj_weight"], G[a + ".in_proj_bias"]
        if self._trains(a + ".in_proj_weight"):
            with self._aside(dq, dkv):
                k.colsum(dq, gb[:D])
                k.colsum(dkv, gb[D:])
                k.linear_wgrad(dq, q_in, gw[:D])
                if kmem is vmem:
                    k.linear_wgrad(dkv, kmem, gw[D:])
                else:
                    k.linear_wgrad(dkv[:, :D], kmem, gw[D:2 * D])
                    k.linear_wgrad(dkv[:, D:], vmem, gw[2 * D:])
        if kmem is vmem:
            dmem = k.linear_dgrad(dkv, wi[D:], residual=dmem)
        else:
            dmem = k.linear_dgrad(dkv[:, :D], wi[D:2 * D], residual=dmem)
            dmem = k.linear_dgrad(dkv[:, D:], wi[2 * D:], residual=dmem)
        if pos_grad is None:
            dx = k.linear_dgrad(dq, wi[:D], residual=dpre)
        else:
            dq_in = k.linear_dgrad(dq, wi[:D])
            k.batch_reduce(dq_in, pos_grad, B, Sq)
            dx = k.add(dpre, dq_in, out=dq_in)
        return dx, dmem

    def _ffn_fwd(self, w1, w2, ln, x, eps, act=RELU, p_out=None, save=True):
        """y = LN(x + drop(W2 drop_h(act(W1 x + b1)) + b2)).  Train mode: the ReLU networks (DETR, text decoder) drop
        the hidden activation too (transformer.py:158); the GELU ones (ViLBERT / BERT intermediate) only the output."""
        W, Pm = self.W, self.P
        if act == RELU and x.shape[1] == 256 and self.fused_layers and x.shape[0] >= 48 * 128:
            # (below ~48 row tiles the three-launch path spreads over more SMs: 3200 decoder rows are 25 tiles)
            # one kernel for the whole sub-layer (csrc/layer_umma.cu): the hidden activation never returns as a GEMM operand
            y, h, pre, st = k.mlp_block_fwd(x, W[w1 + ".weight"], Pm[w1 + ".bias"], W[w2 + ".weight"], Pm[w2 + ".bias"],
                                            Pm[ln + ".weight"], Pm[ln + ".bias"], eps, save=save,
                                            drop_h=self._drop(w1 + ".hidden"), drop_o=self._drop(ln + ".in", p_out))
            return y, (x, h, h, pre, st)
        if act == GELU:
            hpre = torch.empty((x.shape[0], W[w1 + ".weight"].shape[0]), device=self.dev, dtype=BF16)
            h = k.linear(x, W[w1 + ".weight"], Pm[w1 + ".bias"], act=GELU, out2=hpre)
        else:
            h = k.linear(x, W[w1 + ".weight"], Pm[w1 + ".bias"], act=RELU, drop=self._drop(w1 + ".hidden"), drop_mode=k.DROP_POST_ACT)
            hpre = h
        pre = k.linear(h, W[w2 + ".weight"], Pm[w2 + ".bias"], residual=x, drop=self._drop(ln + ".in", p_out),
                       drop_mode=k.DROP_PRE_RESIDUAL)
        y, st = k.layernorm_fwd(pre, Pm[ln + ".weight"], Pm[ln + ".bias"], eps)
        return y, (x, h, hpre, pre, st)

    def _ffn_bwd(self, w1, w2, ln, dy, sv, act=RELU, p_out=None):
        Pm, G = self.P, self.G
        x, h, hpre, pre, st = sv
        dpre, dpre_m = self._ln_bwd(dy, pre, st, Pm[ln + ".weight"], G[ln + ".weight"], G[ln + ".bias"], self._drop(ln + ".in", p_out))
        dh_drop = self._drop(w1 + ".hidden") if act != GELU else None
        # hidden dropout: the saved h is already masked, so relu'(h) (*) mask = [h > 0]; only the 1/(1-p) scale remains
        alpha = dh_drop.scale if dh_drop is not None else 1.0
        if act == RELU and x.shape[1] == 256 and self.fused_bwd and x.shape[0] >= 48 * 128 and (w1 + ".weight.T") in self.W:
            # both data gradients in one tcgen05 launch (the mirror of mlp_block_fwd); the weight / bias gradients follow on the lanes
            self._lin_bwd(w2, h, dpre_m, need_dx=False)          # (issued first: it does not depend on the fused kernel)
            dx, dh = k.mlp_block_bwd(dpre_m, self.W[w2 + ".weight.T"], self.W[w1 + ".weight.T"], h, dpre, alpha=alpha)
            self._lin_bwd(w1, x, dh, need_dx=False)
            return dx
        dh = self._lin_bwd(w2, h, dpre_m, aux=hpre, aux_mode=GRAD_GELU if act == GELU else MASK_RELU, alpha=alpha)
        return self._lin_bwd(w1, x, dh, residual=dpre)

    # ================================================================================================ BERT (no grad)
    def _bert_fwd(self, ids):
        """bert.py:11-22 / HF BertModel in eval mode; ids [B,T] int64 on device -> [B*T,768] bf16."""
        P, W = self.P, self.W
        b = "bert.model"
        B, T = ids.shape
        # key-padding mask: BertTokenizer(padding=True) pads with [PAD] = 0 and hands BertModel attention_mask = (ids != 0)
        # (bert.py:12-21), so [PAD] keys are masked in all 12 self-attention layers; the PAD *rows* still flow on into the
        # co-attention, which is unmasked in the reference (gpv.py:150-154)
        kmask = torch.empty((B, T), device=self.dev, dtype=torch.uint8)
        e = k.gather_rows(P[f"{b}.embeddings.word_embeddings.weight"], ids.reshape(-1), pos=P[f"{b}.embeddings.position_embeddings.weight"],
                          cst=P[f"{b}.embeddings.token_type_embeddings.weight"], T=T, pad_mask=kmask, pad_id=0)
        # bert.py:11-22 runs under no_grad but in the module's train/eval mode: HF's dropouts (embeddings, attention
        # probabilities, both dense outputs) are active while training
        x, _ = k.layernorm_fwd(e, P[f"{b}.embeddings.LayerNorm.weight"], P[f"{b}.embeddings.LayerNorm.bias"], 1e-12, need_stats=False,
                               drop=self._drop(f"{b}.embeddings.out"))
        D = x.shape[1]
        for i in range(12):
            p = f"{b}.encoder.layer.{i}"
            qkv = k.linear(x, W[p + ".qkv"], self.Bcat[p + ".qkv"])
            o, _ = k.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B=B, H=12, Sq=T, Sk=T, dh=64, scale=0.125, need_lse=False,
                                   key_mask=kmask, drop=self._drop(p + ".attention.probs"))
            pre = k.linear(o, W[p + ".attention.output.dense.weight"], P[p + ".attention.output.dense.bias"], residual=x,
                           drop=self._drop(p + ".attention.output.in"), drop_mode=k.DROP_PRE_RESIDUAL)
            x1, _ = k.layernorm_fwd(pre, P[p + ".attention.output.LayerNorm.weight"], P[p + ".attention.output.LayerNorm.bias"], 1e-12,
                                    need_stats=False)
            h = k.linear(x1, W[p + ".intermediate.dense.weight"], P[p + ".intermediate.dense.bias"], act=GELU)
            pre = k.linear(h, W[p + ".output.dense.weight"], P[p + ".output.dense.bias"], residual=x1,
                           drop=self._drop(p + ".output.in"), drop_mode=k.DROP_PRE_RESIDUAL)
            x, _ = k.layernorm_fwd(pre, P[p + ".output.LayerNorm.weight"], P[p + ".output.LayerNorm.bias"], 1e-12, need_stats=False)
        return x

    # ================================================================================================ co-attention
    def _coatt_fwd(self, p, lang, vis, B, Tl, Q):
        """vilbert.py:872-900 with tensor1 = language, tensor2 = vision (gpv.py:149-154)."""
        W, Pm = self.W, self.P
        D, H = self.D, self.h_co
        dh = D // H
        sc = 1.0 / math.sqrt(dh)
        pd = self.p_drop          # vilbert.py:720-727, 833-851: "1" modules (language stream here) use the v_* probabilities
        with self._aside(lang, lane=1):
            qkv1 = k.linear(lang, W[p + ".qkv1"], self.Bcat[p + ".qkv1"])
        qkv2 = k.linear(vis, W[p + ".qkv2"], self.Bcat[p + ".qkv2"])
        self._join(1)
        with self._aside(lang, qkv2, lane=1):                 # language stream (attends to vision) beside the vision stream
            ctx2, lse2 = k.attention_fwd(qkv1[:, :D], qkv2[:, D:2 * D], qkv2[:, 2 * D:], B=B, H=H, Sq=Tl, Sk=Q, dh=dh, scale=sc,
                                         drop=self._drop(p + ".probs2", pd["co_att_v"]))
            pa1 = k.linear(ctx2, W[p + ".biOutput.dense1.weight"], Pm[p + ".biOutput.dense1.bias"], residual=lang,
                           drop=self._drop(p + ".biOutput.LayerNorm1.in", pd["co_hid_l"]), drop_mode=k.DROP_PRE_RESIDUAL)
            att1, sa1 = k.layernorm_fwd(pa1, Pm[p + ".biOutput.LayerNorm1.weight"], Pm[p + ".biOutput.LayerNorm1.bias"], 1e-12)
            o1, f1 = self._ffn_fwd(p + ".v_intermediate.dense", p + ".v_output.dense", p + ".v_output.LayerNorm", att1, 1e-12, GELU,
                                   p_out=pd["co_hid_l"])
        ctx1, lse1 = k.attention_fwd(qkv2[:, :D], qkv1[:, D:2 * D], qkv1[:, 2 * D:], B=B, H=H, Sq=Q, Sk=Tl, dh=dh, scale=sc,
                                     drop=self._drop(p + ".probs1", pd["co_att_l"]))
        pa2 = k.linear(ctx1, W[p + ".biOutput.dense2.weight"], Pm[p + ".biOutput.dense2.bias"], residual=vis,
                       drop=self._drop(p + ".biOutput.LayerNorm2.in", pd["co_hid_v"]), drop_mode=k.DROP_PRE_RESIDUAL)
        att2, sa2 = k.layernorm_fwd(pa2, Pm[p + ".biOutput.LayerNorm2.weight"], Pm[p + ".biOutput.LayerNorm2.bias"], 1e-12)
        o2, f2 = self._ffn_fwd(p + ".t_intermediate.dense", p + ".t_output.dense", p + ".t_output.LayerNorm", att2, 1e-12, GELU,
                               p_out=pd["co_hid_v"])
        self._join(1)
        return o1, o2, (lang, vis, qkv1, qkv2, ctx1, lse1, ctx2, lse2, pa1, sa1, pa2, sa2, f1, f2)

    def _coatt_bwd(self, p, do1, do2, sv, B, Tl, Q, need_dlang=True):
        W, Pm, G = self.W, self.P, self.G
        D, H = self.D, self.h_co
        dh = D // H
        sc = 1.0 / math.sqrt(dh)
        lang, vis, qkv1, qkv2, ctx1, lse1, ctx2, lse2, pa1, sa1, pa2, sa2, f1, f2 = sv
        pd = self.p_drop
        with self._aside(do1, lane=1):                        # language stream beside the vision stream
            datt1 = self._ffn_bwd(p + ".v_intermediate.dense", p + ".v_output.dense", p + ".v_output.LayerNorm", do1, f1, GELU,
                                  p_out=pd["co_hid_l"])
            dpa1, dpa1_m = self._ln_bwd(datt1, pa1, sa1, Pm[p + ".biOutput.LayerNorm1.weight"], G[p + ".biOutput.LayerNorm1.weight"],
                                        G[p + ".biOutput.LayerNorm1.bias"], self._drop(p + ".biOutput.LayerNorm1.in", pd["co_hid_l"]))
            dctx2 = self._lin_bwd(p + ".biOutput.dense1", ctx2, dpa1_m)
        datt2 = self._ffn_bwd(p + ".t_intermediate.dense", p + ".t_output.dense", p + ".t_output.LayerNorm", do2, f2, GELU,
                              p_out=pd["co_hid_v"])
        dpa2, dpa2_m = self._ln_bwd(datt2, pa2, sa2, Pm[p + ".biOutput.LayerNorm2.weight"], G[p + ".biOutput.LayerNorm2.weight"],
                                    G[p + ".biOutput.LayerNorm2.bias"], self._drop(p + ".biOutput.LayerNorm2.in", pd["co_hid_v"]))
        dctx1 = self._lin_bwd(p + ".biOutput.dense2", ctx1, dpa2_m)
        dqkv1, dqkv2 = torch.empty_like(qkv1), torch.empty_like(qkv2)
        self._join(1)
        with self._aside(dqkv1, dqkv2, lane=1):
            k.attention_bwd(qkv1[:, :D], qkv2[:, D:2 * D], qkv2[:, 2 * D:], ctx2, dctx2, lse2, dqkv1[:, :D], dqkv2[:, D:2 * D],
                            dqkv2[:, 2 * D:], B=B, H=H, Sq=Tl, Sk=Q, dh=dh, scale=sc, drop=self._drop(p + ".probs2", pd["co_att_v"]))
        k.attention_bwd(qkv2[:, :D], qkv1[:, D:2 * D], qkv1[:, 2 * D:], ctx1, dctx1, lse1, dqkv2[:, :D], dqkv1[:, D:2 * D],
                        dqkv1[:, 2 * D:], B=B, H=H, Sq=Q, Sk=Tl, dh=dh, scale=sc, drop=self._drop(p + ".probs1", pd["co_att_l"]))
        self._join(1)
        with self._aside(dqkv1, lane=1):
            dlang = self._lin_bwd(p + ".qkv1", lang, dqkv1, residual=dpa1, wkey=p + ".qkv1", need_dx=need_dlang)
        dvis = self._lin_bwd(p + ".qkv2", vis, dqkv2, residual=dpa2, wkey=p + ".qkv2")
        self._join(1)
        return dlang, dvis

    # ================================================================================================ trunk forward
    @torch.no_grad()
    def encode(self, images, qids, save, mask=None):
        """gpv.py:137-175: everything up to `memory`.  Returns a dict of the tensors later stages need.
        mask: [B,H,W] bool padding mask of a mixed-size batch (utils/detr_misc.py:282-299) or None."""
        self.refresh()
        W, Pm = self.W, self.P
        B = images.shape[0]
        d, D, Q = self.d, self.D, self.Q
        s = {"B": B}
        with self._aside(qids, lane=2):                       # BERT (~100 small launches) runs beside the backbone
            qe_b = self._bert_fwd(qids)
            lang = k.linear(qe_b, W["bert_joiner.weight"], Pm["bert_joiner.bias"])
        c5, acts = self._backbone_fwd(images, save)
        _, Hf, Wf, C5 = c5.shape
        S = Hf * Wf
        s.update(acts=acts, c5=c5, S=S, Hf=Hf, Wf=Wf)
        c5f = c5.view(B * S, C5)
        kmask = None
        if mask is None:
            pos, P_rows = self._pos(Hf, Wf), S
        else:
            # backbone.py:76-78 (nearest down-sampling of the mask), position_encoding.py:28-48, key_padding_mask of the
            # encoder self-attention and of the decoder's cross-attention (transformer.py:50-56)
            mf = torch.nn.functional.interpolate(mask[None].float(), size=(Hf, Wf)).to(torch.bool)[0]
            pos, P_rows = k.cast_bf16(sine_position_masked(mf)), B * S
            kmask = mf.reshape(B, S).to(torch.uint8).contiguous()
        s["kmask"] = kmask
        x = k.linear(c5f, W["detr.input_proj.weight"], Pm["detr.input_proj.bias"])
        enc = []
        for i in range(self.n_enc):
            p = f"detr.transformer.encoder.layers.{i}"
            x, sa = self._self_attn_fwd(p, x, pos, P_rows, B, S, self.h_detr, kmask=kmask)
            x, sf = self._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm2", x, 1e-5, save=save)
            enc.append((sa, sf))
        mem = x
        mem_pos = k.add_rowbcast(mem, pos)
        qe = W["detr.query_embed.weight"]
        t = torch.zeros((B * Q, d), device=self.dev, dtype=BF16)
        dec = []
        with self._aside(mem, mem_pos, lane=1):
            kvs = [self._cross_kv(f"detr.transformer.decoder.layers.{i}", mem_pos, mem) for i in range(self.n_dec)]
        for i in range(self.n_dec):
            p = f"detr.transformer.decoder.layers.{i}"
            t, sa = self._self_attn_fwd(p, t, qe, Q, B, Q, self.h_detr)
            if i == 0:
                self._join(1)
            t, sc = self._cross_attn_fwd(p, t, qe, mem_pos, mem, B, Q, S, self.h_detr, kv=kvs[i], kmask=kmask)
            t, sf = self._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm3", t, 1e-5, save=save)
            dec.append((sa, sc, sf))
        # heads (detr_roi_head.py:81-92).  detr_hs = [LN(roi) | hs] is written in place, no cat.
        M = B * Q
        detr_hs = torch.empty((M, C5 + d), device=self.dev, dtype=BF16)
        hs = detr_hs[:, C5:]
        _, st_dn = k.layernorm_fwd(t, Pm["detr.transformer.decoder.norm.weight"], Pm["detr.transformer.decoder.norm.bias"], 1e-5, out=hs)
        lg_detr = torch.zeros((M, 8), device=self.dev, dtype=F32)
        k.linear(hs, W["detr.class_embed.weight"], Pm["detr.class_embed.bias"], out=lg_detr)
        y1 = k.linear(hs, W["detr.bbox_embed.layers.0.weight"], Pm["detr.bbox_embed.layers.0.bias"], act=RELU)
        y2 = k.linear(y1, W["detr.bbox_embed.layers.1.weight"], Pm["detr.bbox_embed.layers.1.bias"], act=RELU)
        boxes = torch.zeros((M, 8), device=self.dev, dtype=F32)
        k.linear(y2, W["detr.bbox_embed.layers.2.weight"], Pm["detr.bbox_embed.layers.2.bias"], act=SIGMOID, out=boxes)
        ldw = (S + 7) // 8 * 8
        wroi = k.roi_weights(boxes, Hf, Wf, ldw)
        roi_raw = torch.empty((M, C5), device=self.dev, dtype=BF16)
        k.gemm(wroi, c5f, roi_raw, M=Q, N=C5, K=S, lda=ldw, ldb=C5, ldd=C5, b_mn=True, batch=B, a_bs=Q * ldw, b_bs=S * C5, d_bs=Q * C5)
        _, st_roi = k.layernorm_fwd(roi_raw, None, None, 1e-5, out=detr_hs[:, :C5])
        vis = k.linear(detr_hs, W["detr_joiner.weight"], Pm["detr_joiner.bias"])
        s["detr_hs_joined"] = vis
        Tl = qids.shape[1]
        self._join(2)
        co = []
        for i in range(self.n_co):
            lang, vis, sv = self._coatt_fwd(f"co_att_transformer.{i}", lang, vis, B, Tl, Q)
            co.append(sv)
        # relevance (gpv.py:162-175): logits = detr logits + Linear(vis); memory = [vis + softmax(logits) tokens | lang]
        logits = torch.zeros((M, 8), device=self.dev, dtype=F32)
        k.linear(vis, W["relevance_predictor.weight"], Pm["relevance_predictor.bias"], residual=lg_detr, out=logits)
        Tm = Q + Tl
        memory = torch.empty((B * Tm, D), device=self.dev, dtype=BF16)
        if self.cfg.relevance_conditioning:
            k.relevance_mix_fwd(vis, logits, Pm["relevance_tokens"], memory, G=Q, out_gstride=Tm, out_off=0)
        else:
            k.copy_rows(vis, memory, M, D, dst_map=(Q, Tm, 0))
        k.copy_rows(lang, memory, B * Tl, D, dst_map=(Tl, Tm, Q))
        s.update(pos=pos, enc=enc, mem=mem, mem_pos=mem_pos, dec=dec, t_final=t, st_dn=st_dn, detr_hs=detr_hs, y1=y1, y2=y2,
                 boxes=boxes, wroi=wroi, ldw=ldw, roi_raw=roi_raw, st_roi=st_roi, qe_b=qe_b, Tl=Tl, co=co, vis=vis, lang=lang,
                 logits=logits, memory=memory, Tm=Tm)
        return s

    @torch.no_grad()
    def decode_text(self, tok_ids, memory, B, Sx, Tm, save):
        """gpv.py:449-466 teacher-forced: tok_ids [B*Sx] int64 -> (logits fp32 [B*Sx, Vp], saved)."""
        W, Pm = self.W, self.P
        emb = k.gather_rows(Pm["answer_input_embedings.embedding_layer.weight"], tok_ids)
        x = k.linear(emb, W["answer_input_embedings.transform.weight"], Pm["answer_input_embedings.transform.bias"])
        layers = []
        with self._aside(memory, lane=1):
            kvs = [self._cross_kv(f"text_decoder.layers.{i}", memory, memory) for i in range(self.n_txt)]
            wc = k.linear(W["answer_head.vocab_embed"], W["answer_head.classifier_transform.weight"], Pm["answer_head.classifier_transform.bias"])
        for i in range(self.n_txt):
            p = f"text_decoder.layers.{i}"
            x, sa = self._self_attn_fwd(p, x, None, 0, B, Sx, self.h_txt, causal=True)
            if i == 0:
                self._join(1)
            x, sc = self._cross_attn_fwd(p, x, None, memory, memory, B, Sx, Tm, self.h_txt, kv=kvs[i])
            x, sf = self._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm3", x, 1e-5)
            layers.append((sa, sc, sf))
        logits = torch.empty((B * Sx, self.Vp), device=self.dev, dtype=F32)
        k.gemm(x, wc, logits, M=B * Sx, N=self.V, K=self.D, lda=x.stride(0), ldb=wc.stride(0), ldd=self.Vp)
        return logits, (emb, layers, x, wc)

    # ================================================================================================ KV-cached decoding
    @torch.no_grad()
    def decode_begin(self, memory, B, Tm, max_len, rep=1):
        """Start autoregressive decoding against `memory` [B*Tm, D] with `rep` hypotheses per sample (beams folded into
        the batch: row b*rep + r).  What the reference recomputes for every generated token (gpv.py:178-196 runs the
        whole decode_text on the growing prefix, i.e. 210 token-decodes for 20 tokens, and the V x 768 x 768 classifier
        transform 20 times) is computed once here: the cross-attention K/V of `memory` for every text-decoder layer and
        the classifier matrix.  Self-attention K/V go to caches [B*rep, max_len, D] that decode_step() appends to."""
        W, Pm, D = self.W, self.P, self.D
        Bp = B * rep
        kv_mem = []
        for i in range(self.n_txt):
            kv_mem.append(self._cross_kv(f"text_decoder.layers.{i}", memory, memory))     # [B*Tm, 2D]: shared by the rep hypotheses of an image
        wc = k.linear(W["answer_head.vocab_embed"], W["answer_head.classifier_transform.weight"], Pm["answer_head.classifier_transform.bias"])
        # one cache [Bp, max_len, 3 D] per layer: the packed q | k | v projection of position t is written in place by ONE GEMM
        qkvc = [torch.empty((Bp, max_len, 3 * D), device=self.dev, dtype=BF16) for _ in range(self.n_txt)]
        return {"Bp": Bp, "rep": rep, "Tm": Tm, "L": max_len, "t": 0, "kv_mem": kv_mem, "wc": wc, "qkvc": qkvc}

    @torch.no_grad()
    def decode_step(self, st, tok_ids):
        """tok_ids [Bp] int64 = the token at position st['t'] of every hypothesis -> logits fp32 [Bp, Vp] of the next
        position.  One query row per hypothesis: 11 launches per layer on [Bp, D] activations."""
        W, Pm, D, H = self.W, self.P, self.D, self.h_txt
        Bp, Tm, L, t = st["Bp"], st["Tm"], st["L"], st["t"]
        assert t < L, "decode_step past max_len"
        dh = D // H
        sc = dh ** -0.5
        emb = k.gather_rows(Pm["answer_input_embedings.embedding_layer.weight"], tok_ids)
        x = k.linear(emb, W["answer_input_embedings.transform.weight"], Pm["answer_input_embedings.transform.bias"])
        for i in range(self.n_txt):
            p = f"text_decoder.layers.{i}"
            wi, bi_ = W[f"{p}.self_attn.in_proj_weight"], Pm[f"{p}.self_attn.in_proj_bias"]
            c = st["qkvc"][i]
            k.linear(x, wi, bi_, out=c[:, t])                             # q | k | v of position t appended in place: row stride L * 3 D
            c2 = c.view(Bp * L, 3 * D)
            # one query row per hypothesis against its own cache: gpvb200_decode_attention (K / V staged once per (hypothesis, head))
            o = k.decode_attention(c[:, t, :D], c2[:, D:2 * D], c2[:, 2 * D:], Bq=Bp, rep=1, H=H, Sk=t + 1, dh=dh, scale=sc,
                                   bs_k=L * 3 * D, bs_v=L * 3 * D)
            pre = k.linear(o, W[f"{p}.self_attn.out_proj.weight"], Pm[f"{p}.self_attn.out_proj.bias"], residual=x)
            x, _ = k.layernorm_fwd(pre, Pm[f"{p}.norm1.weight"], Pm[f"{p}.norm1.bias"], 1e-5, need_stats=False)
            wi, bi_ = W[f"{p}.multihead_attn.in_proj_weight"], Pm[f"{p}.multihead_attn.in_proj_bias"]
            q = k.linear(x, wi[:D], bi_[:D])
            kv = st["kv_mem"][i]
            # the beams of an image share the encoder memory's K / V: staged once per (image, head) for all of them
            o = k.decode_attention(q, kv[:, :D], kv[:, D:], Bq=Bp, rep=st["rep"], H=H, Sk=Tm, dh=dh, scale=sc, bs_k=Tm * 2 * D, bs_v=Tm * 2 * D)
            pre = k.linear(o, W[f"{p}.multihead_attn.out_proj.weight"], Pm[f"{p}.multihead_attn.out_proj.bias"], residual=x)
            x, _ = k.layernorm_fwd(pre, Pm[f"{p}.norm2.weight"], Pm[f"{p}.norm2.bias"], 1e-5, need_stats=False)
            h = k.linear(x, W[f"{p}.linear1.weight"], Pm[f"{p}.linear1.bias"], act=RELU)
            pre = k.linear(h, W[f"{p}.linear2.weight"], Pm[f"{p}.linear2.bias"], residual=x)
            x, _ = k.layernorm_fwd(pre, Pm[f"{p}.norm3.weight"], Pm[f"{p}.norm3.bias"], 1e-5, need_stats=False)
        wc = st["wc"]
        logits = torch.empty((Bp, self.Vp), device=self.dev, dtype=F32)
        k.gemm(x, wc, logits, M=Bp, N=self.V, K=D, lda=x.stride(0), ldb=wc.stride(0), ldd=self.Vp)
        st["t"] = t + 1
        return logits

    @torch.no_grad()
    def decode_reorder(self, st, parent):
        """Beam search: hypothesis r continues hypothesis parent[r] (int64 [Bp]) -> permute the self-attention caches."""
        n = st["t"] * 3 * self.D                               # only the positions decoded so far move
        if "qkvc_alt" not in st:
            st["qkvc_alt"] = [torch.empty_like(c) for c in st["qkvc"]]
        for src, dst in zip(st["qkvc"], st["qkvc_alt"]):
            k.reorder_rows(src, dst, parent, n)
        st["qkvc"], st["qkvc_alt"] = st["qkvc_alt"], st["qkvc"]

    # ================================================================================================ training step
    @torch.no_grad()
    def forward_train(self, images, qids, ans_ids, tgt, mask=None):
        """images [B,3,H,W] fp32, qids [B,Tl] int64, ans_ids [B,S] int64 (device).  tgt: HostTargets (see gpv.py).
        Returns (total_loss fp32 [1] on device, outputs dict)."""
        if self.train_mode:
            self.drop_seed.add_(1)                            # new masks every step; backward re-reads the same value
        # the gradient arenas are cleared here, on a lane beside the forward pass (456 MB of stores that used to open the backward's
        # critical path); the previous step's gradients have been consumed by then (all-reduce and optimizer follow backward)
        with self._aside(lane=3):
            self.grad_arena.zero_()
            self.grad_pack.zero_()
        self._arena_clean = True
        s = self.encode(images, qids, save=True, mask=mask)
        B, Q, D = s["B"], self.Q, self.D
        M = B * Q
        S = ans_ids.shape[1]
        # ---- criterion.  The localisation branch (matcher cost -> assignment -> set criterion: a serial chain of ~75 us that needs only
        # the relevance logits and boxes) runs on a lane beside the text decoder; the answer cross-entropy follows the decoder.
        loss_terms = torch.zeros(4, device=self.dev, dtype=F32)        # [answer CE (weighted), loss_ce, loss_bbox, loss_giou]
        dlg = torch.zeros((M, 8), device=self.dev, dtype=F32)
        dbox = torch.zeros((M, 8), device=self.dev, dtype=BF16)
        idx_q = idx_t = None
        if tgt.n_loc > 0:
            with self._aside(loss_terms, dlg, dbox, s["logits"], s["boxes"], lane=4):
                lg3, bx3 = s["logits"].view(B, Q, 8), s["boxes"].view(B, Q, 8)
                if tgt.Tmax > 0:
                    cost = k.matcher_cost(lg3, bx3, tgt.boxes, tgt.labels, tgt.offsets, tgt.Tmax, *self.cost_w, C=2)
                    idx_q, idx_t = k.lsap(cost, tgt.offsets)
                k.set_criterion(s["logits"], s["boxes"], tgt.boxes, tgt.offsets, idx_q, idx_t, tgt.loc_valid, eos_coef=self.eos_coef,
                                weight_sum=tgt.weight_sum, num_boxes=tgt.num_boxes, wt_ce=self.loss_wts["loss_ce"],
                                wt_bbox=self.loss_wts["loss_bbox"], wt_giou=self.loss_wts["loss_giou"], out3=loss_terms[1:4], dlogits=dlg,
                                dbox_pre=dbox)
        logits_v, sv_txt = self.decode_text(ans_ids.reshape(-1), s["memory"], B, S, s["Tm"], save=True)
        dlogits_v = torch.empty((B * S, self.Vp), device=self.dev, dtype=BF16)
        if self.Vp != self.V:
            dlogits_v.zero_()
        k.ce_fwd_bwd(logits_v[:, :self.V], tgt.ce_targets, tgt.ce_row_weight, loss_terms[0:1], dlogits_v[:, :self.V])
        self._join(4)
        loss = (loss_terms * (self._wts_loc if tgt.n_loc else self._wts_noloc)).sum().reshape(1)
        # (n_loc > 0 always holds for a captured step: the set criterion then yields zeros when no image has boxes)
        s.update(S_ans=S, sv_txt=sv_txt, dlogits_v=dlogits_v, dlg=dlg, dbox=dbox, loss_terms=loss_terms, idx_q=idx_q, idx_t=idx_t)
        self.saved = s
        self._join(3)
        return loss, s

    @torch.no_grad()
    def backward(self):
        """Gradients of the last forward_train() into the arena (self.G).  Consumes the saved activations."""
        s = self.saved
        assert s is not None, "backward() without a forward_train()"
        self.saved = None
        W, Pm, G = self.W, self.P, self.G
        if not getattr(self, "_arena_clean", False):          # (forward_train clears the arenas beside the forward pass)
            self.grad_arena.zero_()
            self.grad_pack.zero_()
        self._arena_clean = False
        B, Q, D, d = s["B"], self.Q, self.D, self.d
        M, S, Tl, Tm, Sx = B * Q, s["S"], s["Tl"], s["Tm"], s["S_ans"]
        depth = self._backward_depth()
        self.last_stage = {4: N_STAGES - 1, 3: 3}.get(depth, 2)   # gradient stages past it hold frozen parameters only: never touched
        # ---- answer head + text decoder
        emb, layers, xf, wc = s["sv_txt"]
        dlv = s["dlogits_v"]
        with self._aside(dlv, xf):                            # classifier-transform branch: beside the text decoder's backward
            dwc = torch.empty((self.V, D), device=self.dev, dtype=BF16)
            k.gemm(dlv, xf, dwc, M=self.V, N=D, K=B * Sx, lda=dlv.stride(0), ldb=xf.stride(0), ldd=D, a_mn=True, b_mn=True)
            self._lin_bwd("answer_head.classifier_transform", W["answer_head.vocab_embed"], dwc, need_dx=False)
        dx = k.gemm(dlv, wc, torch.empty((B * Sx, D), device=self.dev, dtype=BF16), M=B * Sx, N=D, K=self.V, lda=dlv.stride(0),
                    ldb=wc.stride(0), ldd=D, b_mn=True)
        dmemory = None
        for i in range(self.n_txt - 1, -1, -1):
            p = f"text_decoder.layers.{i}"
            sa, sc, sf = layers[i]
            dx = self._ffn_bwd(p + ".linear1", p + ".linear2", p + ".norm3", dx, sf)
            dx, dmemory = self._cross_attn_bwd(p, dx, sc, s["memory"], s["memory"], dmemory, None, B, Sx, Tm, self.h_txt)
            dx = self._self_attn_bwd(p, dx, sa, None, B, Sx, self.h_txt, causal=True, has_pos=False)
            self._join_layer()
        self._lin_bwd("answer_input_embedings.transform", emb, dx, need_dx=False)
        self._done(0)
        # ---- memory split, relevance conditioning
        dvis = torch.empty((M, D), device=self.dev, dtype=BF16)
        dlang = torch.empty((B * Tl, D), device=self.dev, dtype=BF16)
        k.copy_rows(dmemory, dvis, M, D, src_map=(Q, Tm, 0))
        k.copy_rows(dmemory, dlang, B * Tl, D, src_map=(Tl, Tm, Q))
        dlg = s["dlg"]
        if self.cfg.relevance_conditioning:
            k.relevance_mix_bwd(dmemory, s["logits"], Pm["relevance_tokens"], dlg, G["relevance_tokens"], M=M, G=Q, gstride=Tm, off=0)
        dlg_b = k.cast_bf16(dlg)                                             # [M,8] bf16, columns 2.. are zero
        dvis = self._lin_bwd("relevance_predictor", s["vis"], dlg_b[:, :2], residual=dvis)
        # ---- co-attention
        for i in range(self.n_co - 1, -1, -1):
            dlang, dvis = self._coatt_bwd(f"co_att_transformer.{i}", dlang, dvis, s["co"][i], B, Tl, Q)
            self._join_layer()
        self._lin_bwd("bert_joiner", s["qe_b"], dlang, need_dx=False)
        self._done(1)
        # ---- detr_joiner, ROI head, box / class heads
        detr_hs = s["detr_hs"]
        C5 = detr_hs.shape[1] - d
        if depth == 0:                                      # all of DETR frozen: the data gradient stops at its output
            self._lin_bwd("detr_joiner", detr_hs, dvis, need_dx=False)
            self._done(2)
            return self._backward_end()
        d_hs_all = self._lin_bwd("detr_joiner", detr_hs, dvis)                 # [M, C5 + d]
        hs = detr_hs[:, C5:]
        dy2 = self._lin_bwd("detr.bbox_embed.layers.2", s["y2"], s["dbox"][:, :4], aux=s["y2"], aux_mode=MASK_RELU)
        dy1 = self._lin_bwd("detr.bbox_embed.layers.1", s["y1"], dy2, aux=s["y1"], aux_mode=MASK_RELU)
        dhs = self._lin_bwd("detr.bbox_embed.layers.0", hs, dy1, residual=d_hs_all[:, C5:])
        dhs = self._lin_bwd("detr.class_embed", hs, dlg_b[:, :2], residual=dhs)
        droi = k.layernorm_bwd(d_hs_all[:, :C5], s["roi_raw"], s["st_roi"], None, None, None)
        dc5 = torch.empty((B * S, C5), device=self.dev, dtype=BF16)
        k.gemm(s["wroi"], droi, dc5, M=S, N=C5, K=Q, lda=s["ldw"], ldb=C5, ldd=C5, a_mn=True, b_mn=True, batch=B,
               a_bs=Q * s["ldw"], b_bs=Q * C5, d_bs=S * C5)
        dt = k.layernorm_bwd(dhs, s["t_final"], s["st_dn"], Pm["detr.transformer.decoder.norm.weight"],
                             G["detr.transformer.decoder.norm.weight"], G["detr.transformer.decoder.norm.bias"])
        if depth == 1:
            self._done(2)
            return self._backward_end()
        # ---- DETR decoder / encoder
        gq = G["detr.query_embed.weight"]
        dmem = None
        for i in range(self.n_dec - 1, -1, -1):
            p = f"detr.transformer.decoder.layers.{i}"
            sa, sc, sf = s["dec"][i]
            dt = self._ffn_bwd(p + ".linear1", p + ".linear2", p + ".norm3", dt, sf)
            dt, dmem = self._cross_attn_bwd(p, dt, sc, s["mem_pos"], s["mem"], dmem, gq, B, Q, S, self.h_detr, kmask=s["kmask"])
            dt = self._self_attn_bwd(p, dt, sa, gq, B, Q, self.h_detr)
            self._join_layer()
        self._done(2)
        if depth == 2:
            return self._backward_end()
        dx = dmem
        for i in range(self.n_enc - 1, -1, -1):
            p = f"detr.transformer.encoder.layers.{i}"
            sa, sf = s["enc"][i]
            dx = self._ffn_bwd(p + ".linear1", p + ".linear2", p + ".norm2", dx, sf)
            dx = self._self_attn_bwd(p, dx, sa, None, B, S, self.h_detr, kmask=s["kmask"])
            self._join_layer()
        # ---- input_proj: dC5 = (dx Wip + dC5_roi) * relu'(c5)  -> masked gradient of the last bottleneck
        c5f = s["c5"].view(B * S, C5)
        dpre = self._lin_bwd("detr.input_proj", c5f, dx, residual=dc5, aux=c5f, aux_mode=MASK_RELU)
        self._done(3)
        if depth == 3:
            return self._backward_end()
        self._backbone_bwd(dpre.view(s["c5"].shape), s["acts"])
        return self._backward_end()
