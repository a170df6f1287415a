"""Developer tool (GPU box): time the fused sub-layer kernels of csrc/layer_umma.cu against the launches they replace
(graph-replayed, warm, 20 back-to-back launches each; an L2-sized buffer is rewritten between groups)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpv1_b200 import kernels as k

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def mlp(M, dff=2048, seq=0):
    x = torch.randn(M, 256, device=dev).to(BF)
    w1 = (torch.randn(dff, 256, device=dev) / 16).to(BF)
    w2 = (torch.randn(256, dff, device=dev) / 45).to(BF)
    b1, b2, g, b = (torch.randn(n, device=dev) for n in (dff, 256, 256, 256))
    seed = torch.tensor([3], dtype=torch.int64, device=dev)
    dh, do = k.Drop(seed, 1, 0.1), k.Drop(seed, 2, 0.1)

    def unfused(drop):
        h = k.linear(x, w1, b1, act=k.ACT_RELU, drop=dh if drop else None, drop_mode=k.DROP_POST_ACT)
        pre = k.linear(h, w2, b2, residual=x, drop=do if drop else None, drop_mode=k.DROP_PRE_RESIDUAL)
        k.layernorm_fwd(pre, g, b, 1e-5)

    flop = 4.0 * M * 256 * dff
    for drop in (False, True):
        t0 = timeit(lambda: unfused(drop))
        t1 = timeit(lambda: k.mlp_block_fwd(x, w1, b1, w2, b2, g, b, 1e-5, seq_len=seq, drop_h=dh if drop else None,
                                            drop_o=do if drop else None))
        t2 = timeit(lambda: k.mlp_block_fwd(x, w1, b1, w2, b2, g, b, 1e-5, seq_len=seq, save=False))
        print(f"mlp_block M={M} dff={dff} seq={seq} drop={drop}: unfused (3 launches) {t0:.1f} us | fused {t1:.1f} us = "
              f"{flop / t1 * 1e-6:.0f} TFLOP/s | fused, nothing saved {t2:.1f} us", flush=True)


def attn(B, Sq, Sk, self_attn):
    if self_attn:
        qkv = torch.randn(B * Sq, 768, device=dev).to(BF)
        q, kk, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
    else:
        q = torch.randn(B * Sq, 256, device=dev).to(BF)
        kv = torch.randn(B * Sk, 512, device=dev).to(BF)
        kk, v = kv[:, :256], kv[:, 256:]
    wo = (torch.randn(256, 256, device=dev) / 16).to(BF)
    bo, g, b = (torch.randn(256, device=dev) for _ in range(3))
    x = torch.randn(B * Sq, 256, device=dev).to(BF)
    seed = torch.tensor([3], dtype=torch.int64, device=dev)
    dp, do = k.Drop(seed, 1, 0.1), k.Drop(seed, 2, 0.1)
    sc = 32 ** -0.5

    def unfused(drop):
        o, lse = k.attention_fwd(q, kk, v, B=B, H=8, Sq=Sq, Sk=Sk, dh=32, scale=sc, drop=dp if drop else None)
        pre = k.linear(o, wo, bo, residual=x, drop=do if drop else None, drop_mode=k.DROP_PRE_RESIDUAL)
        k.layernorm_fwd(pre, g, b, 1e-5)

    flop = B * 8 * (4.0 * Sq * Sk * 32) + 2.0 * B * Sq * 256 * 256
    for drop in (False, True):
        t0 = timeit(lambda: unfused(drop))
        ta = timeit(lambda: k.attention_fwd(q, kk, v, B=B, H=8, Sq=Sq, Sk=Sk, dh=32, scale=sc, drop=dp if drop else None))
        t1 = timeit(lambda: k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=sc, wo=wo, bo=bo, x=x, gamma=g, beta=b,
                                             drop_p=dp if drop else None, drop_o=do if drop else None))
        t2 = timeit(lambda: k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=sc, drop_p=dp if drop else None))
        print(f"attn_block B={B} Sq={Sq} Sk={Sk} drop={drop}: mma.sync attention + out-proj + LN (3 launches) {t0:.1f} us (attention alone {ta:.1f}) | "
              f"fused {t1:.1f} us = {flop / t1 * 1e-6:.0f} TFLOP/s | tcgen05 core only {t2:.1f} us", flush=True)


if __name__ == "__main__":
    attn(32, 300, 300, True)
    attn(32, 100, 300, False)
    attn(32, 100, 100, True)
    attn(64, 300, 300, True)
    mlp(9600)
    mlp(9600, seq=300)
    mlp(3200, seq=100)
    mlp(19200, seq=300)
