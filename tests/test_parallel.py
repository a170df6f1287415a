"""Data-parallel gradient exchange on CPU: world_size 2, gloo.  The bucket schedule of parallel.GradSync (one
all-reduce per backward stage over contiguous prefixes of the gradient arena) must leave every rank with the mean."""
import os
import sys

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEngine:
    """The three attributes GradSync uses of model.engine.Engine."""

    def __init__(self, n, stage_end):
        self.grad_arena = torch.zeros(n)
        self.stage_end = stage_end
        self.on_stage_done = None
        self.on_backward_end = None
        self.last_stage = len(stage_end) - 1              # Engine.backward lowers it when the tail of the model is frozen


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpv1_b200.parallel import GradSync
    from gpv1_b200.model.spec import N_STAGES
    ends = [8, 8, 24, 40, 48, 56, 64]
    assert len(ends) == N_STAGES
    eng = _FakeEngine(64, ends)
    sync = GradSync()
    sync.attach(eng)
    torch.manual_seed(rank)
    g = torch.randn(64)
    # backward finishes stage after stage; each completed stage is reduced while later ones are still "computing"
    for st in range(N_STAGES):
        lo = ends[st - 1] if st else 0
        eng.grad_arena[lo:ends[st]] = g[lo:ends[st]]
        eng.on_stage_done(st)
    eng.on_backward_end()
    gathered = [torch.zeros(64) for _ in range(world)]
    dist.all_gather(gathered, g)
    mean = torch.stack(gathered).mean(0)
    ok = torch.allclose(eng.grad_arena, mean, atol=1e-6)
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def _frozen_worker(rank, world, port, out):
    """First training phase (freeze_detr_params): backward stops after gradient stage 2, so only buckets 0..2 are reduced;
    the rest of the arena (frozen parameters: zeros) is never touched by a collective."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpv1_b200.parallel import GradSync
    ends = [8, 8, 24, 40, 48, 56, 64]
    eng = _FakeEngine(64, ends)
    eng.last_stage = 2
    sync = GradSync()
    sync.attach(eng)
    torch.manual_seed(10 + rank)
    g = torch.randn(64)
    g[24:] = 0.0                                         # Engine.backward zeroes the arena and never writes past the cut
    eng.grad_arena[40:] = float(rank + 1)                # sentinel: a reduction over these buckets would average it to 1.5
    for st in range(3):
        lo = ends[st - 1] if st else 0
        eng.grad_arena[lo:ends[st]] = g[lo:ends[st]]
        eng.on_stage_done(st)
    eng.on_backward_end()
    gathered = [torch.zeros(64) for _ in range(world)]
    dist.all_gather(gathered, g)
    mean = torch.stack(gathered).mean(0)
    ok = torch.allclose(eng.grad_arena[:24], mean[:24], atol=1e-6) and bool((eng.grad_arena[40:] == float(rank + 1)).all())
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok" if flag.item() == 1.0 else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("schedule", ["3,4,5,6", "all"])
def test_gradsync_frozen_tail_gloo_world2(tmp_path, monkeypatch, schedule):
    """Both bucket schedules (the default merged one and one bucket per stage, GPVB200_DDP_REDUCE_AT)."""
    monkeypatch.setenv("GPVB200_DDP_REDUCE_AT", schedule)
    out = str(tmp_path / "res.txt")
    port = 29700 + (os.getpid() + len(schedule)) % 250
    mp.spawn(_frozen_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


@pytest.mark.parametrize("schedule", ["3,4,5,6", "all"])
def test_gradsync_gloo_world2(tmp_path, monkeypatch, schedule):
    monkeypatch.setenv("GPVB200_DDP_REDUCE_AT", schedule)
    out = str(tmp_path / "res.txt")
    port = 29500 + (os.getpid() + len(schedule)) % 1000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_grad_stage_covers_every_live_parameter():
    sys.path.insert(0, ROOT)
    from gpv1_b200.config import load_config
    from gpv1_b200.model.spec import N_STAGES, gpv_specs, grad_stage, never_gets_grad
    specs = gpv_specs(load_config().model, 512)
    live = [s.name for s in specs if s.kind == "param" and not never_gets_grad(s.name)]
    stages = {grad_stage(n) for n in live}
    assert stages == set(range(N_STAGES))


def _nccl_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import json
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gpv1_b200.config import load_config
    from gpv1_b200.model import GPV
    from gpv1_b200.parallel import GradSync, broadcast_parameters
    from oracle import torch_oracle as TO
    from oracle.make_golden import make_inputs
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "gpv_specs.json")))
    V = g["V"]
    P = TO.make_state([tuple(s) for s in g["specs"]], seed=0)
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(V - 4)]
    model = GPV(load_config().model, vocab=vocab, vocab_embed=P["answer_head.vocab_embed"].numpy())
    model.load_state_dict(P, strict=True)
    model.to(dev)
    model.eval()
    broadcast_parameters(model)
    images, qids, ans, targets = make_inputs(2, 192, 256, 6, 5, 100 + rank, ["CocoCaptioning", "CocoVqa"])
    dt = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in t.items()} for t in targets]
    batch = (images.to(dev), qids.to(dev), ans.to(dev), dt)
    names = ["detr_joiner.weight", "text_decoder.layers.0.linear1.weight", "detr.backbone.0.body.layer2.0.conv2.weight",
             "detr.transformer.encoder.layers.3.self_attn.in_proj_weight", "co_att_transformer.1.biattention.value2.weight"]
    params = dict(model.named_parameters())

    def step():
        for p in model.parameters():
            p.grad = None
        model(*batch).backward()
        torch.cuda.synchronize()
        return {n: params[n].grad.clone() for n in names}

    local = step()                                            # no GradSync attached: rank-local gradients
    mean = {}
    for n in names:
        parts = [torch.zeros_like(local[n]) for _ in range(world)]
        dist.all_gather(parts, local[n])
        mean[n] = torch.stack(parts).mean(0)
    GradSync(model)
    ok = True
    eager = step()
    for n in names:
        ok &= bool((eager[n] - mean[n]).norm() <= 2e-3 * mean[n].norm() + 1e-7)
    model.capture_step(*batch)
    graphed = step()
    for n in names:
        ok &= bool((graphed[n] - mean[n]).norm() <= 2e-3 * mean[n].norm() + 1e-7)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok" if flag.item() == 1.0 else "mismatch")
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gradsync_nccl_two_gpus(tmp_path):
    """On a box with >= 2 GPUs: the stage-bucketed all-reduce (eager and under the captured CUDA graphs) leaves every
    rank with the mean of the rank-local gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "res.txt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
