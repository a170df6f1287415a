"""Developer tool (GPU box): BASELINE.json configs[3] -- beam_size=5 text decode, batch=64, inference only, and the
greedy decode of the same batch.  Prints one JSON line per workload (ms per call, samples/s, generated tokens/s)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import V_BENCH, make_batch, vocab_list  # noqa: E402
from gpv1_b200.config import load_config  # noqa: E402
from gpv1_b200.model import GPV  # noqa: E402


def main(B=64, K=5, reps=5):
    dev = torch.device("cuda:0")
    cfg = load_config()
    model = GPV(cfg.model, vocab=vocab_list(V_BENCH), seed=0).to(dev).eval()
    images, qids, _, _ = make_batch(B, seed=4)
    images, qids = images.to(dev), qids.to(dev)
    L = cfg.model.max_text_len
    for name, fn in (("greedy", lambda: model(images, qids, None)), ("beam5", lambda: model.forward_beam_search(images, qids, K))):
        with torch.no_grad():
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            ms = 1e3 * (time.perf_counter() - t0) / reps
        toks = B * (L if name == "greedy" else K * (L - 1))
        print(json.dumps({"workload": f"configs[3] {name}: batch={B}, max_text_len={L}" + (f", beam_size={K}" if name != "greedy" else ""),
                          "ms_per_call": ms, "samples_per_s": B / ms * 1e3, "decoded_tokens_per_s": toks / ms * 1e3,
                          "note": "wall clock around the public API call (encode + KV-cached decode, eager launches, host bookkeeping included)"}))


if __name__ == "__main__":
    main()
