"""Developer tool (GPU box): run a few representative GEMM shapes of the step standalone (for ncu / quick timing)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpv1_b200 import kernels as k  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
SHAPES = [  # (tag, M, N, K, residual, act)
    ("l1.conv3", 614400, 256, 64, True, k.ACT_RELU),
    ("l1.conv1", 614400, 64, 256, False, k.ACT_RELU),
    ("l2.conv3", 153600, 512, 128, True, k.ACT_RELU),
    ("l3.conv3", 38400, 1024, 256, True, k.ACT_RELU),
    ("l3.conv1", 38400, 256, 1024, False, k.ACT_RELU),
    ("enc.ffn1", 9600, 2048, 256, False, k.ACT_RELU),
    ("enc.ffn2", 9600, 256, 2048, True, k.ACT_NONE),
    ("enc.proj", 9600, 256, 256, True, k.ACT_NONE),
    ("txt.qkv", 640, 2304, 768, False, k.ACT_NONE),
    ("txt.ffn", 640, 768, 768, True, k.ACT_NONE),
]


def main(reps=3, only=None):
    torch.manual_seed(0)
    for tag, M, N, K, res, act in SHAPES:
        if only and tag not in only:
            continue
        x = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
        b = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).to(BF) if res else None
        y = torch.empty(M, N, device=dev, dtype=BF)
        for _ in range(2):
            k.linear(x, w, b, act=act, residual=r, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            k.linear(x, w, b, act=act, residual=r, out=y)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        byts = 2 * (M * K + N * K + M * N * (2 if res else 1))
        fl = 2.0 * M * N * K
        print(f"{tag:10s} M{M} N{N} K{K}: {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s  {fl / us / 1e6:7.1f} TFLOP/s")
        # dgrad + wgrad of the same layer
        dy = torch.randn(M, N, device=dev).to(BF)
        dx = torch.empty(M, K, device=dev, dtype=BF)
        dw = torch.zeros(N, K, device=dev)
        for fn, name in ((lambda: k.linear_dgrad(dy, w, out=dx), "dgrad"), (lambda: k.linear_wgrad(dy, x, dw), "wgrad")):
            fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / reps
            print(f"   {name}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    main(a.reps, a.only)
