"""Host logic of the entry-point mirrors (gpv1_b200/train.py, gpv1_b200/inference.py): learning-rate schedules,
checkpoint key mapping, output decoding, config overrides.  CPU tests use stand-in modules; the GPU tests run the real
loop for a few synthetic iterations, including the frozen first phase and a checkpoint round trip."""
import os
import types

import numpy as np
import pytest
import torch


def test_lr_schedules_match_torch_and_warmup_linear():
    from gpv1_b200.train import lr_multiplier, multistep_multiplier
    # WarmupLinearSchedule(warmup_steps=100, t_total=1000) of pytorch_transformers (train_distr.py:296-302)
    assert lr_multiplier(0, 1000, 0.1) == 0.0
    assert abs(lr_multiplier(50, 1000, 0.1) - 0.5) < 1e-12
    assert lr_multiplier(100, 1000, 0.1) == 1.0
    assert abs(lr_multiplier(550, 1000, 0.1) - 0.5) < 1e-12
    assert lr_multiplier(1000, 1000, 0.1) == 0.0 and lr_multiplier(1200, 1000, 0.1) == 0.0
    # MultiStepLR against torch itself
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, [10, 15, 20, 25, 30, 35], 0.5)
    for epoch in range(40):
        assert abs(opt.param_groups[0]["lr"] - multistep_multiplier(epoch, [10, 15, 20, 25, 30, 35], 0.5)) < 1e-12
        opt.step()
        sch.step()


def test_config_overrides_and_training_block():
    from gpv1_b200.config import load_config
    cfg = load_config(overrides=["training.freeze=True", "training.lr=3e-4", "inputs.query=What is this?", "ckpt=/tmp/x.pth"])
    assert cfg.training.freeze is True and cfg.training.lr == 3e-4 and cfg.training.lr_backbone == 1e-5
    assert cfg.training.lr_milestones == [10, 15, 20, 25, 30, 35] and cfg.training.clip_max_norm == 0.1
    assert cfg.inputs.query == "What is this?" and cfg.ckpt == "/tmp/x.pth"


def test_checkpoint_roundtrip_uses_module_prefix(tmp_path):
    """train_distr.py:382-394 / inference.py:57-62: 'model' keys carry DDP's `module.` prefix; resume ignores tensors whose
    shape changed (train_distr.py:270-274)."""
    from gpv1_b200.train import load_checkpoint, save_checkpoint
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    opt = types.SimpleNamespace(state_dict=lambda: {"state": {}, "t": 7}, load_state_dict=lambda sd: setattr(opt, "loaded", sd["t"]))
    path = str(tmp_path / "model.pth")
    save_checkpoint(path, m, opt, epoch=3, step=42)
    raw = torch.load(path)
    assert all(k.startswith("module.") for k in raw["model"]) and raw["epoch"] == 3 and raw["step"] == 42
    m2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 5))      # second layer changed shape
    before = m2[1].weight.clone()
    epoch, step = load_checkpoint(path, m2, opt)
    assert (epoch, step) == (3, 42) and opt.loaded == 7
    assert torch.equal(m2[0].weight, m[0].weight) and torch.equal(m2[1].weight, before)


def test_decode_outputs_sorts_boxes_and_cuts_at_stop():
    """inference.py:24-49."""
    from gpv1_b200.inference import decode_outputs
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__", "a", "red", "bus"]
    model = types.SimpleNamespace(token_ids_to_words=lambda ids: [[vocab[j] for j in row] for row in ids])
    logits = torch.full((1, 1, 5, len(vocab)), -5.0)
    for t, w in enumerate([4, 5, 6, 2, 6]):
        logits[0, 0, t, w] = 5.0
    rel = torch.tensor([[[0.0, 1.0], [3.0, 0.0], [1.0, 0.0]]])
    boxes = torch.tensor([[[0.1] * 4, [0.2] * 4, [0.3] * 4]])
    out = decode_outputs({"pred_relevance_logits": rel, "pred_boxes": boxes, "answer_logits": logits}, model)[0]
    assert out["answer"] == "a red bus"
    assert np.allclose(out["boxes"][:, 0], [0.2, 0.3, 0.1]) and out["relevance"][0] > out["relevance"][1] > out["relevance"][2]


@pytest.mark.gpu
def test_train_loop_two_phases_and_resume(tmp_path):
    """The loop of train_distr.py on synthetic batches: phase one with the DETR-initialised parameters frozen (no
    gradient kernels, no optimizer state for them, their values unchanged), checkpoint, phase two resumed from it."""
    from gpv1_b200.config import load_config
    from gpv1_b200 import train as T
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ov = ["training.batch_size=2", "training.frozen_batch_size=2", "training.num_epochs=1", "training.frozen_epochs=1",
          "training.log_step=1", f"ckpt_dir={tmp_path}", "training.synthetic_iters=3"]
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(60)]
    logs, built = [], []
    real_gpv = T.GPV

    def small_gpv(cfg_model, vocab=None, vocab_embed=None):
        m = real_gpv(cfg_model, vocab=vocab, vocab_embed=vocab_embed, seed=0)
        m.init_detr_params = [n for n, _ in m.named_parameters() if n.startswith("detr.transformer.")]   # as load_pretr_detr would
        built.append(m)
        return m

    T.GPV = small_gpv
    real_data = T.SyntheticBatches
    T.SyntheticBatches = lambda n, bs, H, W, vocab, seed=0: real_data(n, bs, 160, 192, vocab, seed=seed, Tl=6)
    try:
        cfg = load_config(overrides=ov + ["training.freeze=True"])
        loss1 = T.train(cfg, vocab=vocab, log=logs.append)
        # the loop replays CUDA graphs: every batch shape it met was captured on first sight (training.cuda_graphs = 4 shapes at most)
        assert built[0].auto_capture == 4 and 1 <= len(built[0]._captures) <= 3 and all(c.launches_per_step > 0 for c in built[0]._captures)
        ck = torch.load(os.path.join(str(tmp_path), "model.pth"))
        assert any(k.startswith("module.detr.") for k in ck["model"]) and ck["epoch"] == 0 and ck["step"] == 3
        # the optimizer entry is torch.optim.AdamW's layout: index-keyed state in the reference's four-group parameter order
        from gpv1_b200.optim import group_of
        names = [n for n, _ in small_gpv(cfg.model, vocab=vocab).named_parameters()]
        order = [n for g in range(4) for n in names if group_of(n) == g]
        state = ck["optimizer"]["state"]
        assert [len(g["params"]) for g in ck["optimizer"]["param_groups"]] == [sum(1 for n in names if group_of(n) == g) for g in range(4)]
        frozen_state = [n for i, n in enumerate(order) if n.startswith("detr.transformer.") and i in state]
        assert not frozen_state, frozen_state[:3]
        assert any(i in state for i, n in enumerate(order) if n.startswith("text_decoder."))
        assert all(float(v["step"]) == 3.0 for v in state.values()) and "warmup_scheduler" in ck and "lr" in ck
        cfg2 = load_config(overrides=ov + [f"training.ckpt={os.path.join(str(tmp_path), 'model.pth')}", "training.num_epochs=2"])
        loss2 = T.train(cfg2, vocab=vocab, log=logs.append)
    finally:
        T.GPV, T.SyntheticBatches = real_gpv, real_data
    assert loss1 is not None and loss2 is not None and np.isfinite(loss1) and np.isfinite(loss2)
    assert any("Loading checkpoint at the end of epoch 0" in l for l in logs)
    assert sum("total_loss" in l for l in logs) == 6


@pytest.mark.gpu
def test_device_prefetcher_yields_the_same_batches():
    """data.DevicePrefetcher (double-buffered H2D on a copy stream) hands GPV.forward device tensors equal to the host
    batches, in order, across buffer reuse."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpv1_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = []
    for i in range(5):
        imgs = torch.randn(2, 3, 32, 48, generator=g)
        host.append((imgs, torch.randint(0, 100, (2, 6), generator=g), [{"task": "CocoVqa", "boxes": torch.rand(3, 4, generator=g)} for _ in range(2)]))
    for move in (True, False):      # targets moved here, or left on the host for GPV.forward's packed staging copy (the default)
        got = []
        for imgs, q, tg in DevicePrefetcher(host, "cuda:0", move_targets=move):
            assert imgs.is_cuda and q.is_cuda and tg[0]["boxes"].is_cuda == move and tg[0]["task"] == "CocoVqa"
            got.append((imgs.clone(), q.clone(), tg[1]["boxes"].clone()))       # consumed on the compute stream, like a step would
        torch.cuda.synchronize()
        assert len(got) == 5
        for (hi, hq, ht), (di, dq, db) in zip(host, got):
            assert torch.equal(hi, di.cpu()) and torch.equal(hq, dq.cpu()) and torch.equal(ht[1]["boxes"], db.cpu())
