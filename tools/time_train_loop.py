"""Developer tool (GPU box): steady-state time per iteration of gpv1_b200.train.train() on synthetic batches of the bench shape
(B = 32, 3x480x640, pinned fp32 host images through DevicePrefetcher, clip + AdamW every step), with the loop's CUDA-graph replay
(training.cuda_graphs=4, the default) and with eager launches (training.cuda_graphs=0)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gpv1_b200 import train as T  # noqa: E402
from gpv1_b200.config import load_config  # noqa: E402

vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(8188)]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 24
# four pre-generated batches with pinned pixels, cycled (generating 29 M normals per batch on the host would hide the GPU)
pool = [(i.pin_memory(), q.pin_memory(), t) for i, q, t in T.SyntheticBatches(4, 32, 480, 640, vocab, seed=1000)]
for graphs in (4, 0):
    cfg = load_config(overrides=["training.batch_size=32", "training.num_epochs=1", "training.log_step=1",
                                 f"training.synthetic_iters={iters}", f"training.cuda_graphs={graphs}"])
    stamps = []

    def log(line):
        stamps.append(time.perf_counter())

    t0 = time.perf_counter()
    loss = T.train(cfg, data=[pool[i % 4] for i in range(iters)], vocab=vocab, log=log)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tail = stamps[len(stamps) // 2:]                     # second half of the epoch: every shape has been captured by then
    ms = 1e3 * (tail[-1] - tail[0]) / max(1, len(tail) - 1)
    print(f"training.cuda_graphs={graphs}: {ms:.2f} ms per iteration in steady state ({32e3 / ms:.0f} samples/s), "
          f"epoch of {iters} iterations {wall:.1f} s incl. model build, last loss {loss:.4f}", flush=True)
