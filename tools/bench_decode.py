"""Developer tool (GPU box): BASELINE.json configs[3] -- beam_size=5 text decode, batch=64, inference only -- the
greedy decode of the same batch, and configs[0] (single image, 6-token query, greedy), each eager and as one CUDA graph.  Prints one JSON line per workload (ms per call, samples/s, generated tokens/s)."""
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


def main(K=5, reps=5):
    dev = torch.device("cuda:0")
    cfg = load_config()
    model = GPV(cfg.model, vocab=vocab_list(V_BENCH), seed=0).to(dev).eval()
    L = cfg.model.max_text_len
    runs = []
    for B in (64, 1):
        images, qids, _, _ = make_batch(B, seed=4)
        images, qids = images.to(dev), qids[:, :6].to(dev) if B == 1 else qids.to(dev)      # configs[0]: a 6-token query
        for graphs in (False, True):
            tag = "CUDA graph" if graphs else "eager launches"
            runs.append((f"greedy B={B} ({tag})", B, graphs, "greedy", lambda i=images, q=qids: model(i, q, None)))
            if B == 64:
                runs.append((f"beam5 B={B} ({tag})", B, graphs, "beam5", lambda i=images, q=qids: model.forward_beam_search(i, q, K)))
    for label, B, graphs, name, fn in runs:
        model.inference_graphs = graphs
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
        print(json.dumps({"workload": f"{label}: max_text_len={L}" + (f", beam_size={K}" if name != "greedy" else ""),
                          "ms_per_call": ms, "samples_per_s": B / ms * 1e3, "decoded_tokens_per_s": toks / ms * 1e3,
                          "note": "wall clock around the public API call (encode + KV-cached decode + host bookkeeping)"}))


if __name__ == "__main__":
    main()
