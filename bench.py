#!/usr/bin/env python
"""Benchmark of the GPV-1 data-parallel training hot path (BASELINE.json: samples/sec, img+query fwd+bwd).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's sm_100a path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                      # one rank per GPU, weak scaling (32 samples / GPU)
    python bench.py --impl reference --steps K --warmup W           # the reference arithmetic (fp32 oracle port) on host cores

Workload (BASELINE.json configs[1]; SURVEY.md 8d "Config 2"): per GPU a batch of 32 synthetic 3x480x640 images,
20-token prompts, 20-token teacher-forced answers (CocoCaptioning), 1-8 target boxes per image, random-init weights of
the GPV-1 architecture (227 M parameters, V = 8192), one step = GPV.forward(images, queries, answer_token_ids, targets)
-> loss.backward() including the Hungarian-matched criterion and, for N > 1, the bucketed gradient all-reduce.
Prints ONE JSON line (see the task contract): `value` with inputs resident in HBM, `e2e` through the public module API
from pinned host buffers with a device->host read of the loss every step.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V_BENCH = 8192
H_IMG, W_IMG, T_L, S_ANS = 480, 640, 20, 20
RESNET = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


# ---------------------------------------------------------------------------------------------------- algorithmic work
def algorithmic_gflop(B, H=H_IMG, W=W_IMG, Tl=T_L, S=S_ANS, V=V_BENCH, Q=100):
    """(forward, forward+backward) GFLOP per SAMPLE, 2 FLOP per MAC, from the layer shapes.
    Backward = data gradient + weight gradient (2x forward) wherever a gradient flows; none for the frozen stem/layer1
    and BERT (gpv.py:142-145); weight gradient only where no upstream data gradient is needed."""
    fwd = bwd = 0.0
    h, w = (H + 1) // 2, (W + 1) // 2
    fwd += 2 * h * w * 64 * 147
    h, w = (h + 1) // 2, (w + 1) // 2
    inp = 64
    for li, (planes, blocks, stride) in enumerate(RESNET, start=1):
        for bi in range(blocks):
            s = stride if bi == 0 else 1
            ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
            convs = [(h * w, planes, inp, 1), (ho * wo, planes, planes, 9), (ho * wo, planes * 4, planes, 1)]
            if bi == 0:
                convs.append((ho * wo, planes * 4, inp, 1))
            for ci, (pix, co, cin, kk) in enumerate(convs):
                f = 2 * pix * co * cin * kk
                fwd += f
                if li >= 2:
                    first = li == 2 and bi == 0 and ci in (0, 3)      # layer2.0 conv1 / downsample: no data gradient
                    bwd += f if first else 2 * f
            h, w, inp = ho, wo, planes * 4
    Sv, d, D = h * w, 256, 768

    def both(f):
        nonlocal fwd, bwd
        fwd += f
        bwd += 2 * f

    both(2 * Sv * 2048 * d)                                                        # input_proj
    both(6 * (4 * 2 * Sv * d * d + 4 * 8 * Sv * Sv * 32 + 2 * 2 * Sv * d * 2048))  # DETR encoder
    both(6 * (4 * 2 * Q * d * d + 4 * 8 * Q * Q * 32 + 2 * 2 * Q * d * d + 2 * 2 * Sv * d * d + 4 * 8 * Q * Sv * 32
              + 2 * 2 * Q * d * 2048))                                             # DETR decoder
    both(2 * Q * d * 2 + 2 * Q * (2 * d * d + d * 4) + 2 * Q * Sv * 2048 + 2 * Q * 2304 * D)   # heads, ROI, detr_joiner
    fwd += 12 * (4 * 2 * Tl * D * D + 4 * 12 * Tl * Tl * 64 + 2 * 2 * Tl * D * 3072)           # BERT (no backward)
    fwd += 2 * Tl * D * D
    bwd += 2 * Tl * D * D                                                          # bert_joiner: weight gradient only
    both(3 * (2 * (Tl + Q) * D * 3 * D + 2 * 4 * 16 * Q * Tl * 48 + 2 * (Tl + Q) * D * D + 2 * 2 * (Tl + Q) * D * 3072))
    both(2 * Q * D * 2)
    Tm = Q + Tl
    fwd += 2 * S * D * D
    bwd += 2 * S * D * D
    both(3 * (4 * 2 * S * D * D + 4 * 8 * S * S * 96 + 2 * 2 * S * D * D + 2 * 2 * Tm * D * D + 4 * 8 * S * Tm * 96 + 2 * 2 * S * D * 2048))
    both(2 * V * D * D / B + 2 * S * D * V)                                        # answer head (classifier once per batch)
    return fwd / 1e9, (fwd + bwd) / 1e9


# ---------------------------------------------------------------------------------------------------- synthetic batch
def make_batch(B, seed, V=V_BENCH):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, H_IMG, W_IMG, generator=g)
    qids = torch.randint(1000, 30000, (B, T_L), generator=g)
    ans = torch.randint(4, V, (B, S_ANS), generator=g)
    ans[:, 0], ans[:, -1] = 1, 2
    targets = []
    for b in range(B):
        nb = int(torch.randint(1, 9, (1,), generator=g))
        cxcy = 0.25 + 0.5 * torch.rand(nb, 2, generator=g)
        wh = 0.05 + 0.3 * torch.rand(nb, 2, generator=g)
        targets.append({"task": "CocoCaptioning", "answer": "x", "boxes": torch.cat((cxcy, wh), -1),
                        "labels": torch.zeros(nb, dtype=torch.long), "answer_token_ids": ans[b, 1:]})
    return images, qids, ans, targets


def make_multitask_batches(n, B, seed, V=V_BENCH):
    """BASELINE.json configs[2] (SURVEY 8d "Config 3"): `n` batches of the multitask stream (gpv1_b200.data.SyntheticMultitask:
    captioning / VQA / detection / classification drawn 0.35 / 0.35 / 0.15 / 0.15), answers encoded as GPV.encode_answers
    does for the synthetic vocabulary (`__cls__ w.. __stop__`, padded with `__pad__` to the batch maximum: gpv.py:401-430)."""
    from gpv1_b200.data import SyntheticMultitask
    vocab = vocab_list(V)
    w2i = {w: i for i, w in enumerate(vocab)}
    out = []
    for images, qids, targets in SyntheticMultitask(n, B, H_IMG, W_IMG, vocab, seed=seed, Tl=T_L):
        rows = [[w2i["__cls__"]] + [w2i[w] for w in t.get("answer", "").split()] + [w2i["__stop__"]] for t in targets]
        S = max(len(r) for r in rows)
        ans = torch.tensor([r + [w2i["__pad__"]] * (S - len(r)) for r in rows], dtype=torch.long)
        for b, t in enumerate(targets):
            t["answer_token_ids"] = ans[b, 1:]
        out.append((images, qids, ans, targets))
    return out


def vocab_list(V):
    return ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(V - 4)]


# ---------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_steps(steps, warmup, B_cpu=4, seed=1, workload="configs[1]"):
    """The reference's arithmetic (fp32 oracle port of GPV.forward + criterion + autograd backward) on the host cores.
    Each step is a bounded sample of the workload: B_cpu images of the same shape instead of 32."""
    from oracle import torch_oracle as TO
    import oracle
    oracle.build()
    from gpv1_b200.config import load_config
    from gpv1_b200.model.spec import gpv_specs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = load_config()
    specs = gpv_specs(cfg.model, V_BENCH)
    P = TO.make_state([(s.name, s.shape, s.kind) for s in specs], seed=0)
    Pg = {n: (t.requires_grad_(True) if s.kind == "param" and not n.startswith("bert.") else t) for (n, t), s in zip(P.items(), specs)}
    images, qids, ans, targets = make_batch(B_cpu, seed) if workload == "configs[1]" else make_multitask_batches(1, B_cpu, seed)[0]
    times = []
    for i in range(warmup + steps):
        for t in Pg.values():
            t.grad = None
        t0 = time.perf_counter()
        loss = TO.gpv_forward(Pg, images, qids, ans, targets)
        loss.backward()
        times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    sps = B_cpu * len(timed) / sum(timed)
    return sps, cores, f"{len(timed)} steps of {B_cpu} samples (3x{H_IMG}x{W_IMG}, Tl={T_L}, S={S_ANS}, V={V_BENCH}) fp32 torch on {cores} host threads", \
        1e3 * sum(timed) / len(timed)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 4))                 # bounded sample: ~3 s per step of 4 samples on 16 host threads
    warmup = max(1, min(args.warmup, 1))
    sps, cores, sample, ms = cpu_reference_steps(steps, warmup, workload=args.workload)
    line = {"impl": "reference", "metric": "samples/sec (img+query fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(32, args.workload)},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def gemm_algorithmic_flop(d):
    """Algorithmic FLOPs (2 per MAC) of one gpvb200_gemm call from its descriptor (include/gpvb200.h): mode 0 plain /
    batched GEMM, mode 1 implicit-GEMM convolution (K channels per tap), mode 2 convolution weight gradient (contraction
    over the pixels).  The stem runs its 7x7x3 = 147-deep contraction as 4 taps x 64 padded channels: counted as 147."""
    if d.mode == 0:
        return 2.0 * d.M * d.N * d.K * max(d.batch, 1)
    pix = float(d.n_img) * d.Ho * d.Wo
    if d.mode == 1:
        depth = 147.0 if (d.ntaps == 4 and d.K == 64 and d.N == 64) else float(d.K) * d.ntaps
        return 2.0 * pix * d.N * depth
    return 2.0 * pix * d.M * d.N * d.ntaps


# position of (B, H, Sq, Sk, dh) in the argument list of each attention entry point (include/gpvb200.h, gpv1_b200/kernels.py)
_ATTENTION_ARGS = {"gpvb200_attention_fwd": 10, "gpvb200_attention_fwd_drop": 10, "gpvb200_attention_fwd_bs": 14,
                   "gpvb200_attention_bwd": 18, "gpvb200_attention_bwd_drop": 18, "gpvb200_attn_block_fwd": 7}


def attention_summary(rows, peak_tflops):
    """BASELINE.json's metric also asks for the attention kernel's share of peak.  rows: (entry point, ctypes args, ms) of one
    traced step.  Algorithmic FLOPs of one launch: 4 B H Sq Sk dh forward (Q K^T and P V), 10 B H Sq Sk dh backward (S
    recomputed, dV, dP, dQ, dK), masks and causality not discounted.  Grouped by direction and head width: d_h = 32 is
    the DETR encoder / decoder, 48 the co-attention, 64 BERT, 96 the text decoder."""
    agg = {}
    for name, a, ms in rows:
        i = _ATTENTION_ARGS.get(name)
        if i is None:
            continue
        B, H, Sq, Sk, dh = (int(getattr(x, "value", x)) for x in a[i:i + 5])
        if not (0 < dh <= 256 and 0 < H <= 64 and B > 0 and Sq > 0 and Sk > 0):
            raise ValueError(f"unexpected attention arguments at {name}: {(B, H, Sq, Sk, dh)}")
        kind = "bwd" if "_bwd" in name else "fwd"
        key = f"{kind}_dh{dh}"
        flop = (10.0 if kind == "bwd" else 4.0) * B * H * Sq * Sk * dh
        if name == "gpvb200_attn_block_fwd":                  # tcgen05 attention core + out-proj + residual + LayerNorm in one launch
            key = f"fwd_dh{dh}_attn_block_tcgen05_with_out_proj_ln"
            flop += 2.0 * B * Sq * (H * dh) * (H * dh)
        c = agg.setdefault(key, [0, 0.0, 0.0])
        c[0] += 1
        c[1] += flop
        c[2] += ms
    out = {}
    for key, (n, flop, ms) in sorted(agg.items()):
        tf = flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        out[key] = {"launches_per_step": n, "flop_per_launch": flop / n, "avg_launch_us": 1e3 * ms / n, "achieved": tf,
                    "unit": "TFLOP/s", "frac": tf / peak_tflops}
    return out


def trace_gemm_kernel(model, lib, step_fn):
    """Live, in this run: one extra EAGER step (no graph replay, no concurrent lanes) with every C-ABI call bracketed by
    CUDA events on its launching stream (_C._Counting.trace).  Returns the tcgen05 GEMM kernel's launch count, summed
    algorithmic FLOPs and summed launch durations, and the summed duration of every traced launch of the step."""
    eng = model.engine
    cap, model._captured = model._captured, None
    conc, eng.concurrent = eng.concurrent, False
    hooks, eng.on_stage_done, eng.on_backward_end = (eng.on_stage_done, eng.on_backward_end), None, None   # rank-local: no all-reduce
    try:
        step_fn()                                       # eager warm-up (allocator, tensor-map cache)
        torch.cuda.synchronize()
        lib.trace = []
        step_fn()
        torch.cuda.synchronize()
        tr, lib.trace = lib.trace, None
    finally:
        lib.trace = None
        model._captured, eng.concurrent = cap, conc
        eng.on_stage_done, eng.on_backward_end = hooks
    n = 0
    flop = ms_gemm = ms_all = 0.0
    for name, a, e0, e1 in tr:
        t = e0.elapsed_time(e1)
        ms_all += t
        if name == "gpvb200_gemm":
            n += 1
            flop += gemm_algorithmic_flop(a[0]._obj)
            ms_gemm += t
    return n, flop, ms_gemm, ms_all, len(tr), [(name, a, e0.elapsed_time(e1)) for name, a, e0, e1 in tr if name in _ATTENTION_ARGS]


def workload_name(B, workload="configs[1]"):
    if workload == "multitask":
        return (f"configs[2]: batch={B}/GPU CocoCaptioning-shaped synthetic multitask stream (captioning/VQA/detection/classification "
                f"0.35/0.35/0.15/0.15), 3x{H_IMG}x{W_IMG} + {T_L}-tok prompts, answers padded to the batch maximum (S varies per step), "
                f"full fwd+bwd with the task-filtered criterion, V={V_BENCH}")
    return f"configs[1]: batch={B}/GPU synthetic 3x{H_IMG}x{W_IMG} + {T_L}-tok prompts + {S_ANS}-tok answers, full fwd+bwd with SetCriterion, V={V_BENCH}"


# ---------------------------------------------------------------------------------------------------- extra measurements
def torch_eager_gpu(B, dev, steps=3):
    """The practical bar: the SAME arithmetic executed by stock PyTorch-2.11 eager kernels (cuDNN / cuBLAS / ATen) on
    this B200 -- the fp32 restatement of GPV.forward + criterion + autograd backward (oracle/torch_oracle.py, pinned against the
    reference), once in fp32 (torch defaults: TF32 convolutions, fp32 matmuls) and once under torch.autocast(bfloat16), at the
    bench batch.  A baseline measured next to the product, never on its path."""
    from oracle import torch_oracle as TO
    import oracle
    oracle.build()
    from gpv1_b200.config import load_config
    from gpv1_b200.model.spec import gpv_specs
    specs = gpv_specs(load_config().model, V_BENCH)
    P = TO.make_state([(s.name, s.shape, s.kind) for s in specs], seed=0)
    Pg = {n: (t.to(dev).requires_grad_(True) if s.kind == "param" and not n.startswith("bert.") else t.to(dev)) for (n, t), s in zip(P.items(), specs)}
    images, qids, ans, targets = make_batch(B, seed=1000)
    images, qids, ans = images.to(dev), qids.to(dev), ans.to(dev)
    targets = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in t.items()} for t in targets]
    out = {"batch": B, "what": "oracle/torch_oracle.py (fp32 restatement of the reference modules) run by stock torch eager kernels on this GPU, "
                               "fwd + criterion + autograd bwd, CUDA events, 1 warm-up"}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def step():
            for t in Pg.values():
                t.grad = None
            if ctx is None:
                loss = TO.gpv_forward(Pg, images, qids, ans, targets)
            else:
                with ctx:
                    loss = TO.gpv_forward(Pg, images, qids, ans, targets)
            loss.backward()
            return loss
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"value": B / ms * 1e3, "unit": "samples/s", "ms_per_step": ms, "loss": float(loss.detach())}
    for t in Pg.values():
        t.grad = None
    return out


def measure_decode(model, dev, reps=5, K=5):
    """BASELINE configs[0] (one 480x640 image + a 6-token query, greedy, latency) and configs[3] (beam_size = 5, batch = 64,
    inference only) through the public API (GPV.forward / GPV.forward_beam_search), inputs (uint8 NHWC pixels, token ids) in pinned
    host memory, generated token ids / answers read back to the host inside the timed region; whole-call CUDA graph
    (model.inference_graphs)."""
    was_training, graphs = model.training, model.inference_graphs
    model.eval()
    model.inference_graphs = True
    L = model.cfg.max_text_len
    out = {}
    try:
        with torch.no_grad():
            for key, B in (("configs[0] greedy, 1 image", 1), ("configs[3] beam_size=5, batch=64", 64)):
                _, qids, _, _ = make_batch(B, seed=4)
                g8 = torch.Generator().manual_seed(40 + B)       # the loader's raw format: uint8 NHWC pixels (normalisation fused into the stem)
                images = torch.randint(0, 256, (B, H_IMG, W_IMG, 3), generator=g8, dtype=torch.uint8).pin_memory()
                qids = (qids[:, :6] if B == 1 else qids).contiguous().pin_memory()
                if B == 1:
                    fn = lambda: model(images.to(dev, non_blocking=True), qids.to(dev, non_blocking=True), None)["answer_logits"].argmax(-1).cpu()
                else:
                    fn = lambda: model.forward_beam_search(images.to(dev, non_blocking=True), qids.to(dev, non_blocking=True), K)["answers"]
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                toks = B * (L if B == 1 else K * (L - 1))
                out[key] = {"ms_per_call": ms, "samples_per_s": B / ms * 1e3, "decoded_tokens_per_s": toks / ms * 1e3, "max_text_len": L,
                            "h2d_bytes_per_call": images.numel() + qids.numel() * 8}
    finally:
        model.inference_graphs = graphs
        model._inf_graphs.clear()
        model.train(was_training)
    return out


def measure_encoder_layer(model, B, dev, peak_tflops, reps=20):
    """BASELINE's "fused encoder-decoder attention kernel" target, measured live: ONE DETR encoder layer (transformer.py:148-161:
    QKV projections, 8-head attention over 300 tokens, out-proj + residual + LayerNorm, FFN 256-2048-256 + residual + LayerNorm) at
    the bench batch through the engine's own layer functions -- forward (pos add, QK GEMM, V GEMM, attn_block_fwd, mlp_block_fwd: 5
    launches) and backward -- CUDA events around `reps` back-to-back runs, train mode (dropout on).  Algorithmic FLOPs per sample and
    layer: 4 projections 2 S d^2 each + attention 4 H S^2 d_h + FFN 4 S d d_ff; backward 2x."""
    from gpv1_b200 import _C
    eng = model.engine
    S, d, dff, H = (H_IMG // 32) * (W_IMG // 32), 256, 2048, 8
    p = "detr.transformer.encoder.layers.0"
    x = torch.randn(B * S, d, device=dev).to(torch.bfloat16)
    pos = torch.randn(S, d, device=dev).to(torch.bfloat16)
    dy = (0.1 * torch.randn(B * S, d, device=dev)).to(torch.bfloat16)
    flop_fwd = B * (4 * 2 * S * d * d + 4 * H * S * S * (d // H) + 2 * 2 * S * d * dff)
    train = eng.train_mode
    eng.train_mode = model.training
    eng.refresh()
    out = {}
    try:
        def fwd():
            y1, sa = eng._self_attn_fwd(p, x, pos, S, B, S, H)
            y2, sf = eng._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm2", y1, 1e-5)
            return sa, sf

        def fwd_bwd():
            sa, sf = fwd()
            d1 = eng._ffn_bwd(p + ".linear1", p + ".linear2", p + ".norm2", dy, sf)
            eng._self_attn_bwd(p, d1, sa, None, B, S, H)
            eng._join()

        lib = _C.lib()
        for name, fn, flop in (("fwd", fwd, flop_fwd), ("fwd_bwd", fwd_bwd, 3 * flop_fwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            n0 = lib.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / reps
            tf = flop / us * 1e-6
            out[name] = {"us_per_layer": us, "launches": (lib.launches - n0) // reps, "flop": flop, "achieved": tf, "unit": "TFLOP/s",
                         "frac": tf / peak_tflops}
    finally:
        eng.train_mode = train
        eng.grad_arena.zero_()
    out["what"] = (f"one DETR encoder layer, B={B} x S={S} tokens, d=256, 8 heads, d_ff=2048, dropout {'on' if model.training else 'off'}: "
                   "eager launches of the engine's layer functions (tcgen05 attn_block_fwd + mlp_block_fwd forward; GEMM / mma.sync "
                   "attention / LayerNorm kernels backward), frac of the measured bf16 tensor peak")
    return out


def _spread(xs):
    """median / p10 / p90 of the K timed steps, from one CUDA event between consecutive steps on rank 0."""
    if not xs:
        return None
    v = sorted(xs)
    q = lambda f: v[min(len(v) - 1, max(0, int(round(f * (len(v) - 1)))))]
    return {"median": q(0.5), "p10": q(0.1), "p90": q(0.9), "min": v[0], "max": v[-1], "n": len(v)}


def lib_sha16():
    """sha of the library's sources + nvcc flags (build.source_sha16): the .so's bytes differ between builds of the same code."""
    from gpv1_b200 import build
    return build.source_sha16()


# ---------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--workload", default="configs[1]", choices=["configs[1]", "multitask"],
                    help="configs[1] (default, the headline): fixed-shape captioning batch; multitask: BASELINE configs[2], the all.yaml task mix")
    ap.add_argument("--multitask-batches", type=int, default=8, help="distinct batches the multitask workload cycles through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of replaying the captured CUDA graphs")
    ap.add_argument("--profiling", action="store_true", help="under ncu only: allow fewer than 3 warm-up steps, skip the e2e loop")
    ap.add_argument("--breakdown", default=None, help="write a per-kernel time breakdown of one extra step to this file")
    ap.add_argument("--eval-mode", action="store_true", help="time the dropout-free (model.eval()) arithmetic instead of the training mode")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra measurements of the line (multitask, decode, torch eager GPU bar)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not args.profiling:
        args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # as in round 1 (whose N > 1 lines the driver parsed): the version banner still precedes the JSON line on stdout
        dist.init_process_group("nccl", device_id=dev)

    from gpv1_b200 import _C
    from gpv1_b200.config import load_config
    from gpv1_b200.model import GPV
    from gpv1_b200.parallel import GradSync, broadcast_parameters
    cfg = load_config()
    B = args.batch
    model = GPV(cfg.model, vocab=vocab_list(V_BENCH), seed=0).to(dev)
    model.train(not args.eval_mode)       # training mode: every nn.Dropout site of the reference is active (p = 0.1), as in train_distr.py:407
    sync = GradSync(model) if world > 1 else None
    broadcast_parameters(model)

    multitask = args.workload == "multitask"
    if multitask:
        batches = make_multitask_batches(args.multitask_batches, B, seed=1000 + rank)
    else:
        batches = [make_batch(B, seed=1000 + rank)]
    # resident copies (for `value`) and pinned host copies (for `e2e`)
    dev_b = [(i.to(dev), q.to(dev), a.to(dev), [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in t.items()} for t in tg])
             for i, q, a, tg in batches]
    host_b = [(i.pin_memory(), q.pin_memory(), a.pin_memory(), [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in t.items()} for t in tg])
              for i, q, a, tg in batches]
    images, qids, ans, targets = batches[0]
    d_images, d_qids, d_ans, d_targets = dev_b[0]
    h_images, h_qids, h_ans, h_targets = host_b[0]
    turn = [0]                                         # the multitask workload cycles through its batches; configs[1] has one

    def note(msg):
        if os.environ.get("GPV_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    note("model built, parameters broadcast")
    graph_launches = None
    if not args.no_graph and not args.breakdown:
        seen = set()
        for b in dev_b:                                # one captured step per distinct answer length S
            if b[2].shape[1] not in seen:
                seen.add(b[2].shape[1])
                cap = model.capture_step(*b, add=True)
        graph_launches = cap.launches_per_step

    def step_resident():
        turn[0] += 1
        loss = model(*dev_b[turn[0] % len(dev_b)])
        loss.backward()
        return loss

    def step_e2e():
        turn[0] += 1
        loss = model(*host_b[turn[0] % len(host_b)])
        loss.backward()
        return loss.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps, finish=None, per_step=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [e0]
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
            if per_step is not None:                                # one event between steps: the spread of the K timed steps
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        if finish is not None:
            finish()                                                # (the last step's deferred loss read, inside the timed region)
        host_ms[0] = 1e3 * (time.perf_counter() - t0) / steps      # host time to ENQUEUE one step (no sync inside)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if per_step is not None:
            per_step.extend(a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:]))
        return ms.item()

    note("captured" if graph_launches is not None else "eager")
    for _ in range(args.warmup):
        step_resident()
    note("warm-up done")
    lib = _C.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launches
    step_ms = []
    ms = timed(step_resident, args.steps, per_step=step_ms)
    note("timed region done")
    launches = graph_launches if graph_launches is not None else (lib.launches - n0) // args.steps
    host_enqueue_ms = host_ms[0]
    if args.profiling:
        print(json.dumps({"profiling": True, "ms_per_step_under_profiler": ms / args.steps, "launches_per_step": launches}))
        return
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    note("e2e (fp32 sync) done")
    clocks = sampler.stop() if rank == 0 else None
    loss_val = step_e2e()

    # second number: the full training step = forward + backward (+ all-reduce) + clip_grad_norm_ + AdamW +
    # re-derivation of the packed bf16 weights, with the reference's hyper-parameters (train_distr.py:228-253, 414-428)
    from gpv1_b200.optim import ClipAdamW
    opt = ClipAdamW.for_model(model, cfg.training if hasattr(cfg, "training") else None)

    def step_full():
        turn[0] += 1
        loss = model(*dev_b[turn[0] % len(dev_b)])
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step_full()
    ms_full = timed(step_full, args.steps)
    note("full step done")

    # e2e with the input staged one step ahead (gpv1_b200.data.DevicePrefetcher): every step still moves one
    # batch of pinned host pixels to the device inside the timed region, but on a copy stream, under the previous step
    ms_e2e_pf = None
    if not args.no_graph and not args.breakdown and not multitask:
        from gpv1_b200.data import DevicePrefetcher

        def host_batches():
            while True:
                yield h_images, h_qids, h_targets

        pf = DevicePrefetcher(host_batches(), dev)

        def step_e2e_pf():
            imgs, q, tg = next(pf)
            loss = model(imgs, q, h_ans, tg)
            loss.backward()
            return loss.item()

        for _ in range(2):
            step_e2e_pf()
        ms_e2e_pf = timed(step_e2e_pf, args.steps)

    # third number: the same end-to-end step fed with the loader's raw format, uint8 NHWC pixels, whose
    # ToTensor + Normalize (coco_generic_dataset.py:31-32) are folded into the stem's read: a quarter of the H2D bytes
    ms_e2e_u8 = ms_e2e_u8_pf = None
    if not args.no_graph and not args.breakdown and not multitask:
        g8 = torch.Generator().manual_seed(2000 + rank)
        h_u8 = torch.randint(0, 256, (B, H_IMG, W_IMG, 3), generator=g8, dtype=torch.uint8).pin_memory()
        model.capture_step(h_u8.to(dev), d_qids, d_ans, d_targets, add=True)

        def step_e2e_u8():
            loss = model(h_u8, h_qids, h_ans, h_targets)
            loss.backward()
            return loss.item()

        for _ in range(2):
            step_e2e_u8()
        ms_e2e_u8 = timed(step_e2e_u8, args.steps)

        # the headline `e2e`: the loader's raw uint8 batches staged one step ahead by data.DevicePrefetcher -- every timed step
        # still copies one batch of pinned host pixels + ids + targets to the device and reads the loss back to the host
        from gpv1_b200.data import DevicePrefetcher

        def host_batches_u8():
            while True:
                yield h_u8, h_qids, h_targets

        pf8 = DevicePrefetcher(host_batches_u8(), dev)

        def step_e2e_u8_pf():
            imgs, q, tg = next(pf8)
            loss = model(imgs, q, h_ans, tg)
            loss.backward()
            return loss.item()

        for _ in range(2):
            step_e2e_u8_pf()
        ms_e2e_u8_pf_sync = timed(step_e2e_u8_pf, args.steps)
        note("e2e uint8 + prefetch (synchronous loss read) done")

        # the same loop as gpv1_b200.train runs it: every step's loss is copied to pinned host memory behind the step and read
        # one step late (data.LossReader), so the host enqueues step i+1 while step i computes; the last read is inside the region
        from gpv1_b200.data import LossReader
        reader = LossReader()
        e2e_losses = []

        def step_e2e_u8_pf_deferred():
            imgs, q, tg = next(pf8)
            loss = model(imgs, q, h_ans, tg)
            loss.backward()
            v = reader.push(loss)
            if v is not None:
                e2e_losses.append(v)

        for _ in range(2):
            step_e2e_u8_pf_deferred()
        e2e_losses.clear()
        ms_e2e_u8_pf = timed(step_e2e_u8_pf_deferred, args.steps, finish=lambda: e2e_losses.append(reader.flush()))
        assert len(e2e_losses) == args.steps + 1 and all(math.isfinite(v) for v in e2e_losses), e2e_losses   # (+1: the warm-up's last step)
        note("e2e uint8 + prefetch done")

    # DDP check on hardware: after a step every rank must hold the same (averaged) gradient arena
    grads_equal = None
    if world > 1:
        step_resident()
        torch.cuda.synchronize()
        ga = model.engine.grad_arena
        sig = torch.stack((ga.double().sum(), ga.double().square().sum(), ga[::4097].double().abs().sum())).to(dev)
        allsig = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(allsig, sig)
        grads_equal = all(torch.equal(a, allsig[0]) for a in allsig) and bool(torch.isfinite(sig).all()) and sig[1].item() > 0
        note(f"gradient check across ranks: {grads_equal}")

    # BASELINE configs[2]: the multitask stream (answer length varies per step) on the same replicas, resident and end to end
    multitask_line = None
    if not args.no_extras and not multitask and not args.no_graph and not args.breakdown:
        mt = make_multitask_batches(args.multitask_batches, B, seed=3000 + rank)
        mt_dev = [(i.to(dev), q.to(dev), a.to(dev), [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in t.items()} for t in tg])
                  for i, q, a, tg in mt]
        mt_host = [(i.pin_memory(), q.pin_memory(), a.pin_memory(), [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in t.items()} for t in tg])
                   for i, q, a, tg in mt]
        seen = set()
        for b in mt_dev:
            if b[2].shape[1] not in seen:
                seen.add(b[2].shape[1])
                model.capture_step(*b, add=True)
        mturn = [0]
        note(f"multitask: captured answer lengths {sorted(seen)}")

        def step_mt():
            mturn[0] += 1
            loss = model(*mt_dev[mturn[0] % len(mt_dev)])
            loss.backward()
            return loss

        def step_mt_e2e():
            mturn[0] += 1
            loss = model(*mt_host[mturn[0] % len(mt_host)])
            loss.backward()
            return loss.item()

        for _ in range(max(3, len(mt_dev))):
            step_mt()
        ms_mt = timed(step_mt, args.steps)
        for _ in range(2):
            step_mt_e2e()
        ms_mt_e2e = timed(step_mt_e2e, args.steps)
        note("multitask done")
        multitask_line = {"workload": workload_name(B, "multitask"), "value": world * B * args.steps / (ms_mt / 1e3), "unit": "samples/s",
                          "ms_per_step": ms_mt / args.steps, "answer_lengths": sorted(seen),
                          "e2e": {"value": world * B * args.steps / (ms_mt_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_mt_e2e / args.steps,
                                  "what": "pinned fp32 host batches copied inside the call, loss read back"}}
        del mt_dev, mt_host

    breakdown = None
    if rank == 0 and args.breakdown:
        lib.trace = []
        step_resident()
        torch.cuda.synchronize()
        tr, lib.trace = lib.trace, None
        agg = {}
        for name, a, e0, e1 in tr:
            key = name.replace("gpvb200_", "")
            if name == "gpvb200_gemm":
                d = a[0]._obj
                kind = {(0, 0): "fwd", (0, 1): "dgrad", (1, 1): "wgrad", (1, 0): "a_mn"}[(d.a_mn, d.b_mn)]
                if d.mode == 0:
                    key = f"gemm {kind} M{d.M} N{d.N} K{d.K} b{d.batch} s{d.splits}"
                elif d.mode == 1:
                    key = f"conv {kind} {d.n_img}x{d.Ho}x{d.Wo} N{d.N} K{d.K} taps{d.ntaps} st{d.stride}"
                else:
                    key = f"conv wgrad {d.n_img}x{d.Ho}x{d.Wo} M{d.M} N{d.N} taps{d.ntaps} st{d.stride} s{d.splits}"
            t = e0.elapsed_time(e1)
            c = agg.setdefault(key, [0, 0.0])
            c[0] += 1
            c[1] += t
        breakdown = {kk: {"calls": v[0], "ms": round(v[1], 3)} for kk, v in sorted(agg.items(), key=lambda x: -x[1][1])}
        with open(args.breakdown, "w") as f:
            json.dump({"note": "CUDA-event time per C-ABI entry point over one extra step (serialised by the events; shares, not absolutes)",
                       "ms_per_step_untraced": ms / args.steps, "entries": breakdown}, f, indent=1)

    gemm_trace = None
    if not args.breakdown and world == 1:               # N = 1 only: the scaling lines carry the whole-step roofline
        try:
            gemm_trace = trace_gemm_kernel(model, lib, step_resident)
        except Exception as e:                          # the per-kernel breakdown must never cost the bench line
            note(f"kernel trace failed: {e!r}")
            gemm_trace = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    sps = world * B * args.steps / (ms / 1e3)
    sps_e2e = world * B * args.steps / (ms_e2e / 1e3)
    gfs = [algorithmic_gflop(B, S=b[2].shape[1]) for b in batches]
    gf_fwd, gf_all = sum(g[0] for g in gfs) / len(gfs), sum(g[1] for g in gfs) / len(gfs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved = gf_all * B / ms_step              # TFLOP/s per GPU: GFLOP/sample * samples / ms
    traffic = gemm_traffic = None
    traffic_src = "null: no ncu DRAM-byte pass of THIS build of libgpvb200.so is committed (profiles/step_traffic.json carries the source sha of the library it measured)"
    try:                                     # DRAM bytes from the committed ncu pass, used only when it measured this very library
        tj = json.load(open(os.path.join(ROOT, "profiles", "step_traffic.json")))
        if tj.get("lib_sha16") == lib_sha16():
            traffic, gemm_traffic = tj["dram_bytes_per_step"], tj["gemm_kernel"]["dram_bytes_per_launch"]
            traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum of this build (profiles/step_traffic.json: same csrc/ + include/ + nvcc flags)"
    except (OSError, KeyError, ValueError):
        pass
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "1400 (of fallback)"
    step_roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "what": f"whole step as one unit: {gf_all:.1f} algorithmic GFLOP/sample fwd+bwd ({gf_fwd:.1f} fwd) x {B} samples / step "
                             f"time (graph replay); traffic = {traffic_src}; peak = {peak_src}"}
    roofline = step_roofline
    attention = None
    if gemm_trace is not None:
        try:                                     # "attn kernel %peak" of BASELINE.json's metric: attn_fwd / attn_bwd kernels by head width
            attention = attention_summary(gemm_trace[5], peak)
            attention["what"] = ("gpv::attn_fwd_kernel / attn_bwd_kernel (mma.sync, scores on chip) and gpv::attn_block_fwd_kernel (tcgen05, DETR encoder): algorithmic FLOPs per launch / "
                                 "average launch duration in the same traced eager step as `roofline`; frac of the measured bf16 "
                                 "tensor peak; dh32 = DETR encoder / decoder, dh48 = co-attention, dh64 = BERT, dh96 = text decoder")
        except Exception as e:                   # never lose the bench line over the breakdown
            attention = {"error": repr(e)}
    if gemm_trace is not None and gemm_trace[0] > 0 and gemm_trace[2] > 0:
        n_g, flop_g, ms_g, ms_all, n_all = gemm_trace[:5]
        ach_g = flop_g / (ms_g * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "gpv::umma_gemm_kernel<BN,F> (tcgen05 contraction: every nn.Linear / nn.Conv2d forward, data "
                                                  "gradient and weight gradient of the step)",
                    "achieved": ach_g, "peak": peak, "unit": "TFLOP/s", "frac": ach_g / peak, "traffic": gemm_traffic,
                    "launches_per_step": n_g, "flop_per_launch": flop_g / n_g, "avg_launch_us": 1e3 * ms_g / n_g,
                    "share_of_step_kernel_time": ms_g / ms_all,
                    "what": f"dominant kernel: algorithmic FLOPs per launch (from each call's descriptor, 2 per MAC) / average launch "
                            f"duration over the {n_g} GEMM launches of one eager step of this run, CUDA events on the launching stream "
                            f"({n_all} traced launches in the step); traffic = {traffic_src}; peak = {peak_src}"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:        # reported on rank 0 at N = 1 only (torchrun pins OMP threads to 1)
        v, cores, sample, _ = cpu_reference_steps(steps=2, warmup=1, workload=args.workload)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample}
    eager_bar = decode = encdec = None
    if not args.no_extras and world == 1 and not args.breakdown:
        try:
            encdec = measure_encoder_layer(model, B, dev, peak)
        except Exception as e:
            encdec = {"error": repr(e)}
        try:                                           # the extra measurements must never cost the bench line
            decode = measure_decode(model, dev)
        except Exception as e:
            decode = {"error": repr(e)}
        try:
            del opt
            model._captured, model._captures = None, []
            torch.cuda.empty_cache()
            eager_bar = torch_eager_gpu(B, dev)
            eager_bar["speedup_of_value_over_bf16_autocast"] = sps / eager_bar["bf16_autocast"]["value"]
        except Exception as e:
            eager_bar = {"error": repr(e)}
    def batch_bytes(b):
        i, q, a, tg = b
        return i.numel() * 4 + q.numel() * 8 + a.numel() * 8 + sum(v.numel() * v.element_size() for t in tg for v in t.values() if torch.is_tensor(v))

    h2d = sum(batch_bytes(b) for b in batches) // len(batches)
    line = {"metric": "samples/sec (img+query fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init weights, randn images, random token ids)",
            "config": {"workload": workload_name(B, args.workload), "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": graph_launches is not None,
                       "l2": "per-step working set (4 GB of saved activations) exceeds the 126 MB L2; no explicit flush",
                       "dropout": ("off (model.eval(): the parity arithmetic)" if args.eval_mode else
                                   "on: p=0.1 at every nn.Dropout site (counter-based masks fused into the GEMM epilogues, LayerNorm "
                                   "and attention kernels, regenerated in backward)"), "loss": loss_val},
            "clocks": clocks,
            "e2e": ({"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                     "what": "public API call on pinned fp32 NCHW host images, copied synchronously inside the call; loss read back"}
                    if ms_e2e_u8_pf is None else
                    {"value": world * B * args.steps / (ms_e2e_u8_pf / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d - h_images.numel() * 3,
                     "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e_u8_pf / args.steps,
                     "what": "the loop of gpv1_b200.train (step replayed from the captured graphs): GPV.forward + loss.backward fed by data.DevicePrefetcher with the loader's raw "
                             "format (every step copies one batch of pinned uint8 NHWC host pixels, token ids and targets to the device on the copy "
                             "stream, under the previous step; ToTensor + Normalize are fused into the stem's read) and data.LossReader (every step's "
                             "loss is copied to pinned host memory behind the step and read by the host one step late, the last one inside the timed "
                             "region); e2e_sync_read is the same loop with loss.item() after every backward"}),
            "e2e_sync_read": None if ms_e2e_u8_pf is None else {
                "value": world * B * args.steps / (ms_e2e_u8_pf_sync / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_u8_pf_sync / args.steps,
                "h2d_bytes_per_step": h2d - h_images.numel() * 3, "d2h_bytes_per_step": 4,
                "what": "uint8 + DevicePrefetcher with a synchronous loss.item() after every backward (the host enqueues the next step only then)"},
            "e2e_fp32_sync": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                              "what": "the same call on pinned fp32 NCHW host images copied synchronously inside the call (the reference's loop, train_distr.py:401)"},
            "e2e_prefetch": None if ms_e2e_pf is None else {
                "value": world * B * args.steps / (ms_e2e_pf / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_pf / args.steps,
                "h2d_bytes_per_step": h2d, "what": "e2e with the next batch's H2D copy double-buffered on a copy stream (data.DevicePrefetcher)"},
            "e2e_uint8": None if ms_e2e_u8 is None else {
                "value": world * B * args.steps / (ms_e2e_u8 / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_u8 / args.steps,
                "h2d_bytes_per_step": h2d - h_images.numel() * 3, "what": "e2e with uint8 NHWC host images, normalisation fused into the stem"},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
            "step_ms": _spread(step_ms),
            "full_step": {"value": world * B * args.steps / (ms_full / 1e3), "unit": "samples/s", "ms_per_step": ms_full / args.steps,
                          "what": "fwd + bwd (+ all-reduce) + fused clip_grad_norm_/AdamW (2 launches over the gradient arena) + bf16 weight re-pack"},
            "roofline": roofline, "roofline_step": step_roofline, "attention_kernel": attention,
            "cpu_baseline": cpu, "encdec_block": encdec, "multitask": multitask_line, "decode": decode,
            "torch_eager_gpu": eager_bar}
    if sync is not None:
        line["allreduce_bytes_per_step"] = sync.bytes_per_step
        line["grads_equal_across_ranks"] = grads_equal
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
