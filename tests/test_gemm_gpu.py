"""tcgen05 contraction kernel vs. a plain PyTorch fp32 reference of the same op (bf16 inputs, fp32 accumulate).

Tolerance: outputs are bf16-rounded (rel 2^-8), accumulation order differs -> |err| <= 2e-2 * max|ref| + small.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=2e-2, name=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{name}: max err {err:.4g} vs scale {scale:.4g}"


def _bf(*shape, dev, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 128), (256, 256, 256), (300, 256, 2048), (1000, 768, 2304),
                                   (640, 2304, 768), (77, 2, 256), (3200, 4, 256), (130, 72, 200), (9600, 2048, 256)])
def test_linear_fwd(cuda, M, N, K):
    from gpv1_b200 import kernels as k
    torch.manual_seed(0)
    x, w = _bf(M, K, dev=cuda), _bf(N, K, dev=cuda, scale=K ** -0.5)
    b = torch.randn(N, device=cuda)
    y = k.linear(x, w, b)
    _close(y, x.float() @ w.float().t() + b, name="bias")
    y = k.linear(x, w, None, act=k.ACT_RELU)
    _close(y, torch.relu(x.float() @ w.float().t()), name="relu")
    r = _bf(M, N, dev=cuda)
    y32 = k.linear(x, w, b, residual=r, out_dtype=torch.float32)
    _close(y32, x.float() @ w.float().t() + b + r.float(), tol=2e-3, name="res fp32")
    pre = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    y = k.linear(x, w, b, act=k.ACT_GELU, out2=pre)
    ref_pre = x.float() @ w.float().t() + b
    _close(pre, ref_pre, name="gelu pre")
    _close(y, F.gelu(ref_pre), name="gelu")
    y = k.linear(x, w, b, act=k.ACT_SIGMOID, out_dtype=torch.float32)
    _close(y, torch.sigmoid(ref_pre), tol=2e-3, name="sigmoid")


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 128), (1000, 768, 2304), (640, 2048, 768), (300, 256, 2048)])
def test_linear_dgrad(cuda, M, N, K):
    from gpv1_b200 import kernels as k
    torch.manual_seed(1)
    dy, w = _bf(M, N, dev=cuda), _bf(N, K, dev=cuda, scale=N ** -0.5)
    dx = k.linear_dgrad(dy, w)
    _close(dx, dy.float() @ w.float(), name="dgrad")
    h = _bf(M, K, dev=cuda)
    r = _bf(M, K, dev=cuda)
    dx = k.linear_dgrad(dy, w, aux=h, aux_mode=k.AUX_RELU_MASK, residual=r)
    _close(dx, (dy.float() @ w.float() + r.float()) * (h.float() > 0), name="dgrad relu mask")
    dx = k.linear_dgrad(dy, w, aux=h, aux_mode=k.AUX_GELU_GRAD)
    hf = h.float().requires_grad_(True)
    F.gelu(hf).sum().backward()
    _close(dx, (dy.float() @ w.float()) * hf.grad, name="dgrad gelu")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 256, 256), (3200, 768, 2304), (640, 2048, 768), (9600, 256, 2048),
                                   (333, 2, 256)])
def test_linear_wgrad(cuda, M, N, K):
    from gpv1_b200 import kernels as k
    torch.manual_seed(2)
    # TMA needs 16-byte row strides: narrow heads (N=2,4) keep their activations in 8-wide padded buffers
    ldn = (N + 7) // 8 * 8
    dy, x = _bf(M, ldn, dev=cuda)[:, :N], _bf(M, K, dev=cuda)
    dw = torch.zeros(N, K, device=cuda)
    k.linear_wgrad(dy, x, dw)
    ref = dy.float().t() @ x.float()
    _close(dw, ref, tol=2e-3, name="wgrad")
    k.linear_wgrad(dy, x, dw, splits=1)
    _close(dw, 2 * ref, tol=2e-3, name="wgrad accumulate")


def test_batched(cuda):
    """ROI-style batched product: per image W[100,304] (K-major) x F[300,2048] (MN-major B)."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(3)
    Bn, R, P, C = 3, 100, 300, 2048
    Wt = torch.zeros(Bn, R, 304, device=cuda, dtype=torch.bfloat16)
    Wt[:, :, :P] = _bf(Bn, R, P, dev=cuda, scale=0.1)
    Fm = _bf(Bn, P, C, dev=cuda)
    out = torch.empty(Bn, R, C, device=cuda, dtype=torch.float32)
    k.gemm(Wt, Fm, out, M=R, N=C, K=P, lda=304, ldb=C, ldd=C, b_mn=True, batch=Bn, a_bs=R * 304, b_bs=P * C, d_bs=R * C)
    ref = torch.bmm(Wt[:, :, :P].float(), Fm.float())
    _close(out, ref, tol=2e-3, name="batched")


def _conv_ref(x, w, stride, ksize):
    # x NHWC bf16, w [taps, Cout, Cin] bf16 -> NHWC fp32
    Cout, Cin = w.shape[1], w.shape[2]
    w4 = w.float().view(ksize, ksize, Cout, Cin).permute(2, 3, 0, 1).contiguous()
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w4, stride=stride, padding=ksize // 2)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n,H,W,Cin,Cout,ks,stride", [
    (2, 16, 16, 64, 64, 3, 1), (2, 15, 20, 512, 512, 3, 1), (1, 30, 40, 256, 256, 3, 1), (2, 60, 80, 128, 128, 3, 1),
    (2, 30, 40, 64, 128, 1, 1), (2, 30, 40, 256, 512, 1, 2), (2, 30, 40, 128, 128, 3, 2), (1, 60, 80, 128, 128, 3, 2),
    (1, 31, 41, 64, 64, 3, 2)])
def test_conv_fwd(cuda, n, H, W, Cin, Cout, ks, stride):
    from gpv1_b200 import kernels as k
    torch.manual_seed(4)
    x = _bf(n, H, W, Cin, dev=cuda)
    w = _bf(ks * ks, Cout, Cin, dev=cuda, scale=(Cin * ks * ks) ** -0.5)
    b = torch.randn(Cout, device=cuda)
    ref = _conv_ref(x, w, stride, ks) + b
    y = k.conv(x, w, ksize=ks, stride=stride, bias=b, act=k.ACT_RELU)
    assert y.shape == ref.shape
    _close(y, torch.relu(ref), name="conv relu")
    r = _bf(*ref.shape, dev=cuda)
    y = k.conv(x, w, ksize=ks, stride=stride, bias=b, residual=r, act=k.ACT_RELU)
    _close(y, torch.relu(ref + r.float()), name="conv res relu")


@pytest.mark.parametrize("n,H,W,Cin,Cout,ks,stride", [
    (2, 16, 16, 64, 64, 3, 1), (2, 15, 20, 512, 512, 3, 1), (2, 30, 40, 256, 256, 3, 1), (2, 30, 40, 128, 128, 3, 2),
    (2, 30, 40, 64, 128, 1, 1), (2, 30, 40, 256, 512, 1, 2), (1, 60, 80, 128, 128, 3, 2), (1, 31, 41, 64, 128, 3, 2)])
def test_conv_bwd(cuda, n, H, W, Cin, Cout, ks, stride):
    from gpv1_b200 import kernels as k
    from gpv1_b200 import convops
    torch.manual_seed(5)
    x = _bf(n, H, W, Cin, dev=cuda)
    w = _bf(ks * ks, Cout, Cin, dev=cuda, scale=(Cin * ks * ks) ** -0.5)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    w4 = wf.view(ks, ks, Cout, Cin).permute(2, 3, 0, 1)
    y = F.conv2d(xf.permute(0, 3, 1, 2), w4, stride=stride, padding=ks // 2).permute(0, 2, 3, 1)
    dy = _bf(*y.shape, dev=cuda)
    y.backward(dy.float())
    dx = convops.conv_dgrad(dy, w, ksize=ks, stride=stride, in_hw=(H, W))
    _close(dx, xf.grad, name="conv dgrad")
    dw = torch.zeros(ks * ks, Cout, Cin, device=cuda)
    k.conv_wgrad(dy, x, dw, ksize=ks, stride=stride)
    _close(dw, wf.grad, tol=3e-3, name="conv wgrad")


def test_pair_variant_matches_references(cuda):
    """The CTA-pair variant (`cta_group::2`, umma_gemm_kernel<BN, F, true>: two SMs share the B tile) over
    plain / ragged / batched GEMMs, MN-major B and implicit-GEMM convolutions with fused epilogues, in a process of its
    own with the selection threshold lowered (GPVB200_PAIR is read once per process) so that every eligible shape takes it.
    tools/check_pair.py compares each result with a PyTorch fp32 reference and fails if the variant never ran."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GPVB200_PAIR="8")
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "check_pair.py")], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "failures: 0" in res.stdout
