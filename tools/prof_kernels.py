"""Developer tool (GPU box): launch each non-GEMM kernel of the step once at its bench shape (for `ncu --set full`) and
print a CUDA-event timing table (10 launches each replayed from a CUDA graph) when run without a profiler.

    python tools/prof_kernels.py                       # timings
    ncu --set full --clock-control none -o gpurun_out/ncu_kernels python tools/prof_kernels.py --once
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpv1_b200 import kernels as k  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def bf(*s):
    return torch.randn(*s, device=dev).to(BF)


def cases():
    B = 32
    out = []

    def attn(tag, H, Sq, Sk, dh, causal=False):
        D = H * dh
        q, kk, v = bf(B * Sq, D), bf(B * Sk, D), bf(B * Sk, D)
        st = {}

        def fwd():
            st["o"], st["lse"] = k.attention_fwd(q, kk, v, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=dh ** -0.5, causal=causal)

        def bwd():
            do = st.setdefault("do", bf(B * Sq, D))
            dq, dk, dv = torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)
            k.attention_bwd(q, kk, v, st["o"], do, st["lse"], dq, dk, dv, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=dh ** -0.5, causal=causal)

        flops = 4.0 * B * H * Sq * Sk * dh
        out.append((f"attn_fwd {tag}", fwd, flops, None))
        out.append((f"attn_bwd {tag}", bwd, 3.5 * flops, None))

    attn("enc self S=300 dh=32", 8, 300, 300, 32)
    attn("dec cross 100x300 dh=32", 8, 100, 300, 32)
    attn("dec self S=100 dh=32", 8, 100, 100, 32)
    attn("co-att 100x20 dh=48", 16, 100, 20, 48)
    attn("txt self S=20 dh=96 causal", 8, 20, 20, 96, True)
    attn("txt cross 20x120 dh=96", 8, 20, 120, 96)

    for M, D in ((9600, 256), (3200, 768)):
        x, g, b = bf(M, D), torch.randn(D, device=dev), torch.randn(D, device=dev)
        st = {}

        def lnf(x=x, g=g, b=b, st=st):
            st["y"], st["s"] = k.layernorm_fwd(x, g, b, 1e-5)

        def lnb(x=x, g=g, st=st, D=D):
            dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
            k.layernorm_bwd(st["y"], x, st["s"], g, dg, db)

        out.append((f"layernorm_fwd [{M},{D}]", lnf, None, 4.0 * M * D))
        out.append((f"layernorm_bwd [{M},{D}]", lnb, None, 6.0 * M * D))
        dy, acc = bf(M, D), torch.zeros(D, device=dev)
        out.append((f"colsum [{M},{D}]", lambda dy=dy, acc=acc: k.colsum(dy, acc), None, 2.0 * M * D))

    img = torch.randn(B, 3, 480, 640, device=dev)
    out.append(("stem_s2d fp32 NCHW 32x3x480x640", lambda: k.stem_s2d(img), None, img.numel() * 4 + B * 244 * 324 * 32))
    u8 = torch.randint(0, 256, (B, 480, 640, 3), device=dev, dtype=torch.uint8)
    out.append(("stem_s2d uint8 NHWC", lambda: k.stem_s2d(u8), None, u8.numel() + B * 244 * 324 * 32))
    c1 = bf(B, 240, 320, 64)
    out.append(("maxpool3x3s2 [32,240,320,64]", lambda: k.maxpool3x3s2(c1), None, c1.numel() * 2 * 1.25))

    # matcher, BASELINE.json configs[4]: 100 queries x 50 targets x batch 256
    Bm, Q, T = 256, 100, 50
    lg = torch.randn(Bm, Q, 2, device=dev)
    bx = torch.cat([torch.rand(Bm, Q, 2, device=dev), 0.01 + 0.5 * torch.rand(Bm, Q, 2, device=dev)], -1)
    tb = torch.cat([torch.rand(Bm * T, 2, device=dev), 0.01 + 0.5 * torch.rand(Bm * T, 2, device=dev)], -1)
    lab = torch.zeros(Bm * T, dtype=torch.int64, device=dev)
    off = (torch.arange(Bm + 1, device=dev) * T).to(torch.int32)
    st = {}

    def cost():
        st["c"] = k.matcher_cost(lg, bx, tb, lab, off, T, 1.0, 5.0, 2.0)

    out.append(("matcher_cost 256x100x50", cost, None, Bm * (Q * 6 + T * 4 + Q * T) * 4.0))
    out.append(("lsap 256x100x50", lambda: k.lsap(st["c"], off), None, None))

    V = 8192
    logits = torch.randn(640, V, device=dev)
    tgt = torch.randint(0, V, (640,), device=dev)
    rw = torch.ones(640, device=dev)
    ls = torch.zeros(1, device=dev)
    dl = torch.empty(640, V, device=dev, dtype=BF)
    out.append(("ce_fwd_bwd [640,8192]", lambda: k.ce_fwd_bwd(logits, tgt, rw, ls, dl), None, 640 * V * 6.0))
    return out


def main():
    once = "--once" in sys.argv
    torch.manual_seed(0)
    for name, fn, flops, byts in cases():
        fn()
        torch.cuda.synchronize()
        if once:
            continue
        g = torch.cuda.CUDAGraph()           # 10 launches replayed from a graph: no host launch gaps in the timing
        with torch.cuda.graph(g):
            for _ in range(10):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 10
        extra = ""
        if flops:
            extra += f"  {flops / us / 1e6:7.1f} TFLOP/s"
        if byts:
            extra += f"  {byts / us / 1e3:7.1f} GB/s (algorithmic bytes)"
        print(f"{name:36s} {us:8.1f} us{extra}")


if __name__ == "__main__":
    main()
