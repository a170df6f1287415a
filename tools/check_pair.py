"""Developer / test tool (GPU): the CTA-pair variant of the contraction kernel (umma_gemm_kernel<256, F, true>)
against plain PyTorch fp32 references of the same ops, over plain / ragged / batched GEMMs, MN-major B (data gradients) and
implicit-GEMM convolutions with fused epilogues.  The variant is selected inside gpvb200_gemm; this script lowers its
threshold (GPVB200_PAIR, read once per process) so that every eligible shape below takes it, and checks that it did.

    python tools/check_pair.py            # exit code 0 = all shapes match and the pair variant ran
"""
import os
import sys

os.environ.setdefault("GPVB200_PAIR", "8")
os.environ.setdefault("GPVB200_PAIR_BN", "128")     # cover the 128-column pair tiles too
WGRAD = False   # (the weight-gradient extension of the pair variant was removed in round 2: it never ran on hardware)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    from gpv1_b200 import _C, convops
    from gpv1_b200 import kernels as k
    dev = torch.device("cuda:0")
    lib = _C.lib()
    fails = []

    def bf(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)

    def close(name, got, ref, tol=2e-2):
        err = (got.float() - ref.float()).abs().max().item()
        scale = ref.float().abs().max().item() + 1e-6
        n0 = lib.gpvb200_gemm_pair_launches()
        ok = err <= tol * scale
        print(f"{'ok  ' if ok else 'FAIL'} {name}: max err {err:.4g} / scale {scale:.4g}   (pair launches so far {n0})", flush=True)
        if not ok:
            fails.append(name)

    def conv_ref(x, w, stride, ks):
        Cout, Cin = w.shape[1], w.shape[2]
        w4 = w.float().view(ks, ks, Cout, Cin).permute(2, 3, 0, 1).contiguous()
        return F.conv2d(x.float().permute(0, 3, 1, 2), w4, stride=stride, padding=ks // 2).permute(0, 2, 3, 1).contiguous()

    torch.manual_seed(0)
    # ---- plain GEMMs (K-major A and B): even / odd tile counts, ragged M, K not a multiple of 64 or 128, N = 256..2048
    for M, N, K in [(38400, 128, 1152), (4096, 256, 1024), (2100, 512, 520), (9600, 256, 2048), (38400, 256, 1024), (1000, 768, 2304), (640, 2048, 768), (300, 256, 2048)]:
        x, w, b = bf(M, K), bf(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
        ref = x.float() @ w.float().t() + b
        close(f"linear {M}x{N}x{K} bias", k.linear(x, w, b), ref)
        r = bf(M, N)
        close(f"linear {M}x{N}x{K} bias+res+relu", k.linear(x, w, b, residual=r, act=k.ACT_RELU), torch.relu(ref + r.float()))
        close(f"linear {M}x{N}x{K} fp32 out", k.linear(x, w, b, out_dtype=torch.float32), ref, tol=2e-3)
    # ---- data gradients: MN-major B
    for M, N, K in [(38400, 1152, 128), (4096, 1024, 256), (2100, 520, 512), (9600, 2048, 256)]:
        dy, w = bf(M, N), bf(N, K, scale=N ** -0.5)
        h, r = bf(M, K), bf(M, K)
        close(f"dgrad {M}x{N}->{K}", k.linear_dgrad(dy, w), dy.float() @ w.float())
        close(f"dgrad {M}x{N}->{K} mask+res", k.linear_dgrad(dy, w, aux=h, aux_mode=k.AUX_RELU_MASK, residual=r),
              (dy.float() @ w.float() + r.float()) * (h.float() > 0))
    # ---- batched, MN-major B, fp32 output (the ROI product)
    Bn, R, P, C = 3, 300, 600, 2048
    Wt, Fm = bf(Bn, R, P, scale=0.1), bf(Bn, P, C)
    out = torch.empty(Bn, R, C, device=dev, dtype=torch.float32)
    k.gemm(Wt, Fm, out, M=R, N=C, K=P, lda=P, ldb=C, ldd=C, b_mn=True, batch=Bn, a_bs=R * P, b_bs=P * C, d_bs=R * C)
    close("batched 3x300x2048x600 (MN-major B)", out, torch.bmm(Wt.float(), Fm.float()), tol=2e-3)
    # ---- implicit-GEMM convolutions: forward (fused bias / residual / ReLU) and data gradients (stride 1 and 2)
    for n, H, W, Cin, Cout, ks, stride in [(8, 60, 80, 128, 128, 3, 1), (32, 30, 40, 256, 256, 3, 1), (4, 30, 40, 256, 256, 3, 1),
                                           (3, 15, 20, 512, 512, 3, 1), (2, 60, 80, 128, 256, 3, 1),
                                           (2, 30, 40, 1024, 256, 1, 1), (2, 30, 40, 256, 256, 3, 2), (1, 31, 41, 128, 256, 3, 2)]:
        x, w, b = bf(n, H, W, Cin), bf(ks * ks, Cout, Cin, scale=(Cin * ks * ks) ** -0.5), torch.randn(Cout, device=dev)
        ref = conv_ref(x, w, stride, ks) + b
        close(f"conv {n}x{H}x{W} {Cin}->{Cout} k{ks} s{stride} relu", k.conv(x, w, ksize=ks, stride=stride, bias=b, act=k.ACT_RELU), torch.relu(ref))
        r = bf(*ref.shape)
        close(f"conv {n}x{H}x{W} {Cin}->{Cout} k{ks} s{stride} res relu",
              k.conv(x, w, ksize=ks, stride=stride, bias=b, residual=r, act=k.ACT_RELU), torch.relu(ref + r.float()))
    for n, H, W, Cin, Cout, ks, stride in [(8, 60, 80, 128, 128, 3, 1), (32, 30, 40, 256, 256, 3, 1), (4, 30, 40, 256, 256, 3, 1),
                                           (2, 30, 40, 256, 1024, 1, 1), (2, 30, 40, 256, 256, 3, 2), (8, 60, 80, 256, 256, 3, 2)]:
        x, w = bf(n, H, W, Cin), bf(ks * ks, Cout, Cin, scale=(Cin * ks * ks) ** -0.5)
        xf = x.float().requires_grad_(True)
        w4 = w.float().view(ks, ks, Cout, Cin).permute(2, 3, 0, 1)
        y = F.conv2d(xf.permute(0, 3, 1, 2), w4, stride=stride, padding=ks // 2).permute(0, 2, 3, 1)
        dy = bf(*y.shape)
        y.backward(dy.float())
        close(f"conv dgrad {n}x{H}x{W} {Cout}->{Cin} k{ks} s{stride}", convops.conv_dgrad(dy, w, ksize=ks, stride=stride, in_hw=(H, W)), xf.grad)
    if WGRAD:   # weight gradients dW[N,K] += dy[M,N]^T x[M,K]: MN-major A and B, contraction over the rows, split-K with fp32 atomics
        for M, N, K in [(9600, 256, 2048), (38400, 256, 1024), (38400, 1024, 256), (3200, 768, 2304), (9600, 2048, 256), (640, 2048, 768)]:
            dy, x = bf(M, N), bf(M, K)
            dw = torch.zeros(N, K, device=dev)
            n_before = lib.gpvb200_gemm_pair_launches()
            k.linear_wgrad(dy, x, dw)
            ref = dy.float().t() @ x.float()
            close(f"wgrad {M}: {N}x{K} (pair: {lib.gpvb200_gemm_pair_launches() > n_before})", dw, ref, tol=2e-3)
            k.linear_wgrad(dy, x, dw)
            close(f"wgrad {M}: {N}x{K} accumulate", dw, 2 * ref, tol=2e-3)
    torch.cuda.synchronize()
    n_pair = lib.gpvb200_gemm_pair_launches()
    print(f"pair launches: {n_pair}; failures: {len(fails)}")
    if n_pair == 0:
        print("FAIL: the pair variant never ran")
        return 1
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
