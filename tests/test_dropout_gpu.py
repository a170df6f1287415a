"""Train-mode dropout (p = 0.1 at every nn.Dropout site of the reference: transformer.py:155-160, vilbert.py:443-851,
nn.TransformerDecoderLayer, HF BERT).  Masks are counter-based (csrc/common.cuh) and cannot be bit-compatible with
torch's Philox stream, so each fused kernel is checked against plain torch arithmetic / autograd fed with the mask the
kernel itself uses, exported by gpvb200_dropout_mask."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _drop(cuda, site, p=0.1, step=7):
    from gpv1_b200 import kernels as k
    seed = torch.tensor([step], dtype=torch.int64, device=cuda)
    return k.Drop(seed, site, p)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def test_mask_statistics_and_independence(cuda):
    from gpv1_b200 import kernels as k
    m1 = k.dropout_mask(4096, 768, _drop(cuda, 11)).float()
    assert abs(m1.mean().item() - 0.9) < 2e-3
    assert abs(m1.mean(0).min().item() - 0.9) < 0.03 and abs(m1.mean(1).min().item() - 0.9) < 0.06   # no dead rows / columns
    m2 = k.dropout_mask(4096, 768, _drop(cuda, 12)).float()            # another site
    m3 = k.dropout_mask(4096, 768, _drop(cuda, 11, step=8)).float()    # another step
    for other in (m2, m3):
        agree = (m1 == other).float().mean().item()
        assert abs(agree - 0.82) < 5e-3                                # independent masks agree on 0.9^2 + 0.1^2
    assert torch.equal(m1, k.dropout_mask(4096, 768, _drop(cuda, 11)).float())   # reproducible
    odd = k.dropout_mask(33, 301, _drop(cuda, 5)).float()              # odd N: pairs do not straddle rows
    assert abs(odd.mean().item() - 0.9) < 0.02


@pytest.mark.parametrize("M,N,K", [(640, 768, 768), (300, 256, 2048), (9600, 256, 256)])
def test_gemm_dropout_before_residual(cuda, M, N, K):
    """y = res + dropout(x W^T + b)   (transformer.py:155-156: src = src + self.dropout1(src2))."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(1)
    x, w, b, r = torch.randn(M, K, device=cuda).to(BF), (torch.randn(N, K, device=cuda) / K ** 0.5).to(BF), torch.randn(N, device=cuda), \
        torch.randn(M, N, device=cuda).to(BF)
    d = _drop(cuda, 21)
    y = k.linear(x, w, b, residual=r, drop=d, drop_mode=k.DROP_PRE_RESIDUAL)
    mask = k.dropout_mask(M, N, d).float()
    ref = (x.float() @ w.float().t() + b) * mask * d.scale + r.float()
    assert _rel(y, ref) < 6e-3
    zero = (mask == 0)
    assert (y.float()[zero] - r.float()[zero]).abs().max().item() == 0   # dropped elements pass the residual through exactly


def test_gemm_dropout_after_activation(cuda):
    """h = dropout(relu(x W^T + b))   (transformer.py:158: linear2(dropout(activation(linear1(src)))))."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(2)
    M, N, K = 3200, 2048, 256
    x, w, b = torch.randn(M, K, device=cuda).to(BF), (torch.randn(N, K, device=cuda) / K ** 0.5).to(BF), torch.randn(N, device=cuda)
    d = _drop(cuda, 22)
    h = k.linear(x, w, b, act=k.ACT_RELU, drop=d, drop_mode=k.DROP_POST_ACT)
    mask = k.dropout_mask(M, N, d).float()
    ref = torch.relu(x.float() @ w.float().t() + b) * mask * d.scale
    assert _rel(h, ref) < 6e-3
    assert h.float()[mask == 0].abs().max().item() == 0


def test_layernorm_dropout_outputs(cuda):
    from gpv1_b200 import kernels as k
    torch.manual_seed(3)
    M, D = 640, 768
    x, g, b = torch.randn(M, D, device=cuda).to(BF), torch.randn(D, device=cuda), torch.randn(D, device=cuda)
    d = _drop(cuda, 31)
    mask = k.dropout_mask(M, D, d).float()
    y0, st = k.layernorm_fwd(x, g, b, 1e-12)
    y1, _ = k.layernorm_fwd(x, g, b, 1e-12, drop=d)
    assert _rel(y1, y0.float() * mask * d.scale) < 5e-3
    dy = torch.randn(M, D, device=cuda).to(BF)
    dg, db = torch.zeros(D, device=cuda), torch.zeros(D, device=cuda)
    dx0 = k.layernorm_bwd(dy, x, st, g, dg, db)
    dx, dxm = k.layernorm_bwd(dy, x, st, g, torch.zeros_like(dg), torch.zeros_like(db), drop=d)
    assert torch.equal(dx, dx0)
    assert _rel(dxm, dx0.float() * mask * d.scale) < 5e-3
    assert dxm.float()[mask == 0].abs().max().item() == 0


@pytest.mark.parametrize("H,Sq,Sk,dh,causal", [(8, 300, 300, 32, False), (8, 100, 300, 32, False), (16, 100, 20, 48, False),
                                               (8, 20, 20, 96, True), (8, 20, 120, 96, False)])
def test_attention_dropout_forward_backward(cuda, H, Sq, Sk, dh, causal):
    """O = dropout(softmax(scale Q K^T)) V and its gradients against autograd with the kernel's own mask."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(4)
    B, D = 2, H * dh
    q, kk, v = (torch.randn(B * S, D, device=cuda).to(BF) for S in (Sq, Sk, Sk))
    d = _drop(cuda, 41)
    scale = dh ** -0.5
    o, lse = k.attention_fwd(q, kk, v, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=scale, causal=causal, drop=d)
    mask = k.dropout_mask(B * H * Sq, Sk, d).float().view(B, H, Sq, Sk)
    qf, kf, vf = (t.float().view(B, -1, H, dh).transpose(1, 2).requires_grad_(True) for t in (q, kk, v))
    s = (qf @ kf.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(torch.ones(Sq, Sk, dtype=torch.bool, device=cuda).triu(1), float("-inf"))
    ref = ((torch.softmax(s, -1) * mask * d.scale) @ vf)
    ref2 = ref.transpose(1, 2).reshape(B * Sq, D)
    assert _rel(o, ref2) < 1.5e-2
    do = torch.randn(B * Sq, D, device=cuda).to(BF)
    ref2.backward(do.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)
    k.attention_bwd(q, kk, v, o, do, lse, dq, dk, dv, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=scale, causal=causal, drop=d)
    for got, leaf, S in ((dq, qf, Sq), (dk, kf, Sk), (dv, vf, Sk)):
        want = leaf.grad.transpose(1, 2).reshape(B * S, D)
        assert _rel(got, want) < 2.5e-2, (_rel(got, want))
