"""Data-parallel gradient exchange on CPU: world_size 2, gloo.  The bucket schedule of parallel.GradSync (one
all-reduce per backward stage over contiguous prefixes of the gradient arena) must leave every rank with the mean."""
import os
import sys

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


def test_gradsync_gloo_world2(tmp_path):
    out = str(tmp_path / "res.txt")
    port = 29500 + os.getpid() % 1000
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
