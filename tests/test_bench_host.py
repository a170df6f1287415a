"""Host-side logic of bench.py and the synthetic multitask stream (BASELINE.json configs[2]; SURVEY.md 8d "Config 3"):
task mix, answer encoding as GPV.encode_answers does it, algorithmic FLOP accounting.  No GPU, no kernels."""
import collections
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_multitask_stream_follows_the_task_mix():
    from gpv1_b200.data import TASK_MIX, SyntheticMultitask
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(50)]
    assert abs(sum(p for _, p, _, _ in TASK_MIX) - 1.0) < 1e-12
    counts = collections.Counter()
    n = 0
    for imgs, qids, targets in SyntheticMultitask(40, 16, 8, 12, vocab, seed=3, Tl=5):
        assert imgs.shape == (16, 3, 8, 12) and qids.shape == (16, 5) and len(targets) == 16
        for t in targets:
            counts[t["task"]] += 1
            n += 1
            spec = {name: (words, boxes) for name, _, words, boxes in TASK_MIX}[t["task"]]
            if spec[0] is None:
                assert "answer" not in t                                  # detection: encoded as `__cls__ __stop__`
            else:
                assert spec[0][0] <= len(t["answer"].split()) <= spec[0][1]
            if spec[1] is None:
                assert "boxes" not in t                                   # only detection samples carry boxes
            else:
                nb = t["boxes"].shape[0]
                assert spec[1][0] <= nb <= spec[1][1] and t["labels"].shape == (nb,) and t["labels"].dtype == torch.long
                assert (t["boxes"] > 0).all() and (t["boxes"] < 1).all()
    for name, p, _, _ in TASK_MIX:
        assert abs(counts[name] / n - p) < 0.06, (name, counts[name] / n)
    # same seed -> same stream; another seed (another rank) -> another stream
    a = [t["task"] for _, _, tg in SyntheticMultitask(2, 8, 4, 4, vocab, seed=7) for t in tg]
    b = [t["task"] for _, _, tg in SyntheticMultitask(2, 8, 4, 4, vocab, seed=7) for t in tg]
    c = [t["task"] for _, _, tg in SyntheticMultitask(2, 8, 4, 4, vocab, seed=8) for t in tg]
    assert a == b and a != c


def test_multitask_answers_are_encoded_like_encode_answers():
    """bench.make_multitask_batches must produce exactly what GPV.encode_answers (gpv.py:377-430 of the reference) gives
    for the same targets: `__cls__ words __stop__`, `__pad__` up to the batch maximum, targets = ids[:, 1:]."""
    import bench
    from gpv1_b200.model.gpv import GPV
    batches = bench.make_multitask_batches(3, 6, seed=5, V=64)
    vocab = bench.vocab_list(64)
    stub = types.SimpleNamespace(word_to_idx={w: i for i, w in enumerate(vocab)}, vision_token=torch.zeros(1),
                                 cfg=types.SimpleNamespace(answering_type="generation", max_text_len=20))
    for images, qids, ans, targets in batches:
        _, ids = GPV.encode_answers(stub, targets)
        assert torch.equal(ids, ans)
        S = ans.shape[1]
        assert S == max(len(t.get("answer", "").split()) for t in targets) + 2
        for b, t in enumerate(targets):
            assert torch.equal(t["answer_token_ids"], ans[b, 1:])
            if "answer" not in t:
                assert ans[b, 0] == 1 and ans[b, 1] == 2 and (ans[b, 2:] == 0).all()


def test_algorithmic_flop_accounting():
    import bench
    from gpv1_b200._C import GemmDesc
    d = GemmDesc()
    d.mode, d.M, d.N, d.K, d.batch = 0, 128, 256, 512, 3
    assert bench.gemm_algorithmic_flop(d) == 2.0 * 128 * 256 * 512 * 3
    d = GemmDesc()
    d.mode, d.n_img, d.Ho, d.Wo, d.N, d.K, d.ntaps = 1, 2, 30, 40, 128, 128, 9           # 3x3 convolution, 128 -> 128 channels
    assert bench.gemm_algorithmic_flop(d) == 2.0 * 2 * 30 * 40 * 128 * 128 * 9
    d = GemmDesc()
    d.mode, d.n_img, d.Ho, d.Wo, d.N, d.K, d.ntaps = 1, 2, 240, 320, 64, 64, 4            # the stem: 4 taps x 64 padded channels = 7x7x3
    assert bench.gemm_algorithmic_flop(d) == 2.0 * 2 * 240 * 320 * 64 * 147
    d = GemmDesc()
    d.mode, d.n_img, d.Ho, d.Wo, d.M, d.N, d.ntaps = 2, 2, 30, 40, 256, 128, 1            # 1x1 weight gradient
    assert bench.gemm_algorithmic_flop(d) == 2.0 * 2 * 30 * 40 * 256 * 128
    # the closed-form step total the bench's whole-step roofline uses (70.0 fwd / 180.1 fwd+bwd measured with FlopCounterMode)
    fwd, both = bench.algorithmic_gflop(32)
    assert abs(fwd - 69.4) < 0.5 and abs(both - 179.6) < 1.0
    assert bench.algorithmic_gflop(32, S=12)[1] < both


def test_attention_summary_from_trace_rows():
    """bench.attention_summary reads (B, H, Sq, Sk, dh) at the positions the wrappers in gpv1_b200/kernels.py pass them
    (checked here against the wrappers' source) and turns (entry point, args, ms) rows into per-kernel rates."""
    import ctypes
    import inspect
    import re
    import bench
    from gpv1_b200 import kernels
    # the positions in bench._ATTENTION_ARGS must be where `B, H, Sq, Sk, dh` sit in each call of kernels.py
    src = inspect.getsource(kernels)
    for name, pos in bench._ATTENTION_ARGS.items():
        m = re.search(name + r"\((.*?)\),\s*\"att", src, re.S)
        if m is None:
            assert name == "gpvb200_attention_fwd"          # only reached through the _bs form
            continue
        depth, args, cur = 0, [], ""
        for ch in m.group(1):
            if ch == "," and depth == 0:
                args.append(cur.strip())
                cur = ""
                continue
            depth += ch in "([" 
            depth -= ch in ")]"
            cur += ch
        args.append(cur.strip())
        assert args[pos:pos + 5] == ["B", "H", "Sq", "Sk", "dh"], (name, args[pos:pos + 5])
    p0 = ctypes.c_void_p(0)
    i64 = ctypes.c_int64
    fwd = ("gpvb200_attention_fwd_bs", (p0,) * 6 + (i64(768),) * 4 + (i64(0),) * 4 + (32, 8, 300, 300, 32, 0, ctypes.c_float(0.1), p0), 0.035)
    bwd = ("gpvb200_attention_bwd_drop", (p0,) * 10 + (i64(768),) * 8 + (32, 8, 300, 300, 32, 0, ctypes.c_float(0.1), p0, ctypes.c_uint32(1),
                                                                  ctypes.c_float(0.1), p0), 0.087)
    other = ("gpvb200_gemm", (p0, p0), 1.0)
    blk = ("gpvb200_attn_block_fwd", (p0, i64(768)) * 3 + (p0, 32, 8, 300, 300, 32, ctypes.c_float(0.1), p0), 0.044)
    out = bench.attention_summary([fwd, fwd, bwd, other, blk], 1386.5)
    f = 4.0 * 32 * 8 * 300 * 300 * 32
    kb = "fwd_dh32_attn_block_tcgen05_with_out_proj_ln"
    assert set(out) == {"fwd_dh32", "bwd_dh32", kb}
    assert abs(out[kb]["flop_per_launch"] - (f + 2.0 * 32 * 300 * 256 * 256)) < 1
    assert out["fwd_dh32"]["launches_per_step"] == 2 and abs(out["fwd_dh32"]["flop_per_launch"] - f) < 1
    assert abs(out["fwd_dh32"]["achieved"] - f / 35e-6 / 1e12) < 1e-6 * out["fwd_dh32"]["achieved"]
    assert abs(out["bwd_dh32"]["achieved"] - 2.5 * f / 87e-6 / 1e12) < 1e-6 * out["bwd_dh32"]["achieved"]
    assert abs(out["bwd_dh32"]["frac"] - out["bwd_dh32"]["achieved"] / 1386.5) < 1e-12


def test_backward_depth_follows_the_frozen_set():
    """Engine._backward_depth (the frozen-tail cut of the first training phase, train_distr.py:136-140) on the real
    parameter list: which prefix of the DETR sub-graph still needs gradients decides where backward stops."""
    from gpv1_b200.config import load_config
    from gpv1_b200.model.engine import Engine
    from gpv1_b200.model.spec import gpv_specs, never_gets_grad
    specs = gpv_specs(load_config().model, 64)
    live = [s.name for s in specs if s.kind == "param" and not never_gets_grad(s.name)]
    eng = types.SimpleNamespace(G={n: None for n in live}, frozen=set())
    depth = lambda: Engine._backward_depth(eng)
    assert depth() == 4
    detr = [n for n in live if n.startswith("detr.")]
    assert any(n.startswith("detr.backbone.") for n in detr) and "detr.query_embed.weight" in detr
    eng.frozen = set(detr)
    assert depth() == 0                                                  # freeze_detr_params with a full DETR checkpoint
    eng.frozen = set(detr) - {"detr.class_embed.weight", "detr.class_embed.bias"}
    assert depth() == 1                                                  # the 2-way class head does not match DETR's 92 classes
    eng.frozen = {n for n in detr if not n.startswith("detr.transformer.decoder.layers.5.")}
    assert depth() == 2
    eng.frozen = {n for n in detr if n.startswith("detr.backbone.")}
    assert depth() == 3
    eng.frozen = {n for n in detr if n.startswith("detr.backbone.") and ".layer4." not in n}
    assert depth() == 4
    eng.frozen = {n for n in live if not n.startswith("detr.")}           # freezing everything else never cuts the DETR backward
    assert depth() == 4


def test_host_targets_single_copy_staging_equals_piecewise_path():
    """HostTargets under a captured step: host-resident targets are packed into one staging buffer and copied once
    (gpv.py HostTargets._stage_host); the result must equal the piecewise gather used for device-resident targets, for
    mixed tasks, ragged boxes, samples without answers / without boxes, and across reuse of the two staging buffers."""
    import bench
    from gpv1_b200.model.gpv import HostTargets
    loss_wts = {"loss_caption": 5e-2, "loss_vqa": 1.0, "loss_cls": 1.0, "loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}
    B, Q, Tcap = 12, 100, 16
    hw = bench.H_IMG, bench.W_IMG
    bench.H_IMG, bench.W_IMG = 8, 8                    # the images play no role here
    try:
        batches = bench.make_multitask_batches(24, B, seed=9, V=64)
    finally:
        bench.H_IMG, bench.W_IMG = hw
    lengths = collections.Counter(b[2].shape[1] for b in batches)
    S = lengths.most_common(1)[0][0]
    batches = [b for b in batches if b[2].shape[1] == S][:4]
    assert len(batches) >= 3                            # both staging buffers get reused
    fast_static = HostTargets.alloc_static(B, S, Tcap, "cpu")
    slow_static = HostTargets.alloc_static(B, S, Tcap, "cpu")
    for images, qids, ans, targets in batches + batches[:1]:
        f = HostTargets(targets, B, S, Q, loss_wts, 0.1, "cpu", static=fast_static)
        g = HostTargets(targets, B, S, Q, loss_wts, 0.1, "cpu", static=slow_static, fast=False)
        sumT = sum(f.sizes)
        assert f.sizes == g.sizes and f.n_text == g.n_text and f.n_loc == g.n_loc and f.Tmax == g.Tmax == Tcap
        assert torch.equal(f.offsets, g.offsets) and torch.equal(f.loc_valid, g.loc_valid)
        assert torch.equal(f.ce_row_weight, g.ce_row_weight) and torch.equal(f.ce_targets, g.ce_targets)
        assert torch.equal(f.boxes[:sumT], g.boxes[:sumT]) and torch.equal(f.labels[:sumT], g.labels[:sumT])
        # against the batch itself
        assert int(f.offsets[-1]) == sumT == sum(t["boxes"].shape[0] for t in targets if "boxes" in t)
        ce = f.ce_targets.view(B, S)
        for b, t in enumerate(targets):
            if "answer" in t:
                assert torch.equal(ce[b, :S - 1], ans[b, 1:]) and ce[b, S - 1] == 0
            else:
                assert not ce[b].any() and not f.ce_row_weight.view(B, S)[b].any()
    # the un-captured path (fresh tensors per step) gives the same values
    h = HostTargets(targets, B, S, Q, loss_wts, 0.1, "cpu")
    assert torch.equal(h.ce_targets, f.ce_targets) and torch.equal(h.offsets, f.offsets) and torch.equal(h.ce_row_weight, f.ce_row_weight)
    assert torch.equal(h.boxes, f.boxes[:sumT]) and torch.equal(h.labels, f.labels[:sumT])


def test_encode_answers_equals_reference_golden():
    """GPV.encode_answers / token_ids_to_words against what the REFERENCE's own methods (gpv.py:377-441) returned for the same
    targets (tests/golden/gpv_answers.json, written by oracle/make_golden.py `answers` from the unmodified reference): lower-casing,
    `__pad__` up to the batch maximum, truncation to max_text_len, `__unk__` for out-of-vocabulary words, the empty answer."""
    import json
    import os
    from gpv1_b200.model.gpv import GPV
    from oracle.ref_harness import make_vocab
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gpv_answers.json")))
    words, _ = make_vocab(gold["V"], 0)
    stub = types.SimpleNamespace(word_to_idx={w: i for i, w in enumerate(words)}, vocab=list(words), vision_token=torch.zeros(1),
                                 cfg=types.SimpleNamespace(answering_type="generation", max_text_len=gold["max_text_len"]))
    for name, case in gold["cases"].items():
        padded, ids = GPV.encode_answers(stub, case["targets"])
        assert [list(p) for p in padded] == case["padded"], name
        assert ids.tolist() == case["ids"], name
        assert GPV.token_ids_to_words(stub, ids) == case["words"], name
    assert max(len(r) for r in gold["cases"]["long"]["ids"]) == gold["max_text_len"]          # truncation was exercised
    unk = stub.word_to_idx["__unk__"]
    assert any(unk in r for r in gold["cases"]["mixed"]["ids"])                                 # and the OOV path


def test_step_spread_and_source_hash():
    """bench._spread (median / p10 / p90 of the timed steps) and build.source_sha16 (the identity bench.py uses to decide
    whether profiles/step_traffic.json measured the library it is running: a hash of csrc/ + include/ + nvcc flags, not of the .so's
    bytes, which differ between builds of the same code)."""
    import bench
    from gpv1_b200 import build
    sp = bench._spread([float(i) for i in range(1, 21)])
    assert sp["n"] == 20 and sp["min"] == 1.0 and sp["max"] == 20.0 and sp["p10"] <= sp["median"] <= sp["p90"]
    assert sp["median"] in (10.0, 11.0) and sp["p10"] == 3.0 and sp["p90"] == 18.0
    assert bench._spread([]) is None and bench._spread([2.5])["median"] == 2.5
    a, b = build.source_sha16(), build.source_sha16()
    assert a == b and len(a) == 16 and int(a, 16) >= 0
    assert bench.lib_sha16() == a
    flags = list(build.FLAGS)
    try:
        build.FLAGS.append("-DGPV_SOMETHING_ELSE")
        assert build.source_sha16() != a                                  # the flags are part of the identity
    finally:
        build.FLAGS[:] = flags
    assert build.source_sha16() == a
