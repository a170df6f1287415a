"""Developer tool (GPU, under ncu): ONE eager forward_beam_search call (B = 64, beam 5, max_text_len 20) after a warm-up call, so that
`ncu --metrics gpu__time_duration.sum` lists every kernel of the call.  tools/gpu_r3n.sh summarises the list by kernel name."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import H_IMG, V_BENCH, W_IMG, make_batch, vocab_list  # noqa: E402
from gpv1_b200.config import load_config  # noqa: E402
from gpv1_b200.model import GPV  # noqa: E402

dev = torch.device("cuda:0")
model = GPV(load_config().model, vocab=vocab_list(V_BENCH), seed=0).to(dev).eval()
_, qids, _, _ = make_batch(64, seed=4)
images = torch.randint(0, 256, (64, H_IMG, W_IMG, 3), dtype=torch.uint8).to(dev)
qids = qids.to(dev)
with torch.no_grad():
    model.forward_beam_search(images, qids, 5)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model.forward_beam_search(images, qids, 5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(len(out["answers"]))
