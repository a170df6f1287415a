"""Developer tool (GPU box): relative-L2 error of each primitive kernel vs fp32 torch on bf16-representable inputs."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpv1_b200 import kernels as k  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def bfr(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(BF)


torch.manual_seed(0)
M, K, N = 600, 256, 256
x, w, b = bfr(M, K), bfr(N, K, scale=1 / 16), torch.randn(N, device=dev)
print("linear fwd", rel(k.linear(x, w, b), x.float() @ w.float().t() + b))
dy = bfr(M, N, scale=0.1)
print("linear dgrad", rel(k.linear_dgrad(dy, w), dy.float() @ w.float()))
dw = torch.zeros(N, K, device=dev)
k.linear_wgrad(dy, x, dw)
print("linear wgrad", rel(dw, dy.float().t() @ x.float()))
g, be = torch.randn(K, device=dev), torch.randn(K, device=dev)
y, st = k.layernorm_fwd(x, g, be, 1e-5)
xf = x.float().requires_grad_(True)
ref = F.layer_norm(xf, (K,), g, be, 1e-5)
print("ln fwd", rel(y, ref))
dyl = bfr(M, K, scale=0.1)
ref.backward(dyl.float())
dg, db = torch.zeros(K, device=dev), torch.zeros(K, device=dev)
print("ln bwd dx", rel(k.layernorm_bwd(dyl, x, st, g, dg, db), xf.grad))
for (B, H, Sq, Sk, dh, causal, sc) in [(2, 8, 63, 63, 32, False, 1.0), (2, 8, 300, 300, 32, False, 1.0), (3, 8, 100, 42, 32, False, 1.0),
                                        (4, 8, 11, 11, 96, True, 1.0), (4, 8, 11, 106, 96, False, 1.0), (3, 16, 100, 7, 48, False, 1.0),
                                        (2, 8, 300, 300, 32, False, 3.0)]:
    D = H * dh
    q, kk, v = bfr(B * Sq, D, scale=sc), bfr(B * Sk, D, scale=sc), bfr(B * Sk, D)
    o, lse = k.attention_fwd(q, kk, v, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=dh ** -0.5, causal=causal)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, kk, v))
    qh = qf.view(B, Sq, H, dh).transpose(1, 2)
    kh = kf.view(B, Sk, H, dh).transpose(1, 2)
    vh = vf.view(B, Sk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(Sq, Sk, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    r = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B * Sq, D)
    do = bfr(B * Sq, D, scale=0.1)
    r.backward(do.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)
    k.attention_bwd(q, kk, v, o, do, lse, dq, dk, dv, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=dh ** -0.5, causal=causal)
    print(f"attn B{B} H{H} {Sq}x{Sk} dh{dh} causal{int(causal)} sc{sc}: fwd {rel(o, r):.4f} dq {rel(dq, qf.grad):.4f} dk {rel(dk, kf.grad):.4f} dv {rel(dv, vf.grad):.4f}")
