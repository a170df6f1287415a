"""Developer tool (GPU box): step-by-step backward error inside one encoder layer."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpv1_b200 import kernels as k  # noqa: E402
from gpv1_b200.config import load_config  # noqa: E402
from gpv1_b200.model import GPV  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402

cuda = torch.device("cuda:0")
BF = torch.bfloat16
GOLD = os.path.join(ROOT, "tests", "golden")
g = json.load(open(os.path.join(GOLD, "gpv_specs.json")))
V = g["V"]
P = TO.make_state([tuple(s) for s in g["specs"]], seed=0)
vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(V - 4)]
model = GPV(load_config().model, vocab=vocab, vocab_embed=P["answer_head.vocab_embed"].numpy())
model.load_state_dict(P, strict=True)
model.to(cuda)
eng = model.engine
eng.refresh()
Pd = {n: (t.to(cuda).to(BF).float() if t.dtype.is_floating_point and t.dim() >= 2 else t.to(cuda)) for n, t in P.items()}


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


torch.manual_seed(0)
B, S, D = 2, 63, 256
p = "detr.transformer.encoder.layers.2"
x = (torch.randn(B * S, D, device=cuda)).to(BF)
dy = (0.1 * torch.randn(B * S, D, device=cuda)).to(BF)
eng.grad_arena.zero_()
y, sf = eng._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm2", x, 1e-5)
xs, h, hpre, pre, st = sf
Pm, W, G = eng.P, eng.W, eng.G
dpre = k.layernorm_bwd(dy, pre, st, Pm[p + ".norm2.weight"], G[p + ".norm2.weight"], G[p + ".norm2.bias"])
dh = k.linear_dgrad(dpre, W[p + ".linear2.weight"], aux=hpre, aux_mode=k.AUX_RELU_MASK)
dx = k.linear_dgrad(dh, W[p + ".linear1.weight"], residual=dpre)

xf = x.float().requires_grad_(True)
hf = torch.relu(TO.lin(Pd, p + ".linear1", xf))
hf.retain_grad()
pf = xf + TO.lin(Pd, p + ".linear2", hf)
pf.retain_grad()
o = TO.ln(Pd, p + ".norm2", pf, 1e-5)
o.backward(dy.float())
print("fwd h", rel(h, hf), "pre", rel(pre, pf), "y", rel(y, o))
print("bwd dpre", rel(dpre, pf.grad), "dh", rel(dh, hf.grad), "dx", rel(dx, xf.grad))
# same with exact inputs at each stage
dh2 = k.linear_dgrad(pf.grad.to(BF), W[p + ".linear2.weight"], aux=hpre, aux_mode=k.AUX_RELU_MASK)
print("dh from exact dpre", rel(dh2, hf.grad), " mask mismatch frac", ((h.float() > 0) != (hf > 0)).float().mean().item())
print("norms: pre", pf.norm().item(), "x", xf.norm().item(), "dpre", pf.grad.norm().item(), "dx", xf.grad.norm().item(),
      "dgrad part", (xf.grad - pf.grad).norm().item())

print("---- self-attention block")
pos = torch.randn(S, D, device=cuda).to(BF)
eng.grad_arena.zero_()
y, sa = eng._self_attn_fwd(p, x, pos, S, B, S, 8)
xs, qk_in, qkv, o, lse, pre, st = sa
dpre = k.layernorm_bwd(dy, pre, st, Pm[p + ".norm1.weight"], G[p + ".norm1.weight"], G[p + ".norm1.bias"])
do = k.linear_dgrad(dpre, W[p + ".self_attn.out_proj.weight"])
dqkv = torch.empty_like(qkv)
k.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                B=B, H=8, Sq=S, Sk=S, dh=32, scale=32 ** -0.5)
wi = W[p + ".self_attn.in_proj_weight"]
dx1 = k.linear_dgrad(dqkv[:, 2 * D:], wi[2 * D:], residual=dpre)
dx2 = k.linear_dgrad(dqkv[:, :2 * D], wi[:2 * D], residual=dx1)

xf = x.float().view(B, S, D).requires_grad_(True)
qk = xf + pos.float()[None]
Wt, bt = Pd[p + ".self_attn.in_proj_weight"], Pd[p + ".self_attn.in_proj_bias"]
qkvf = torch.cat([torch.nn.functional.linear(qk, Wt[:2 * D], bt[:2 * D]), torch.nn.functional.linear(xf, Wt[2 * D:], bt[2 * D:])], -1)
qkvf.retain_grad()
qh = qkvf[..., :D].view(B, S, 8, 32).transpose(1, 2) * 32 ** -0.5
kh = qkvf[..., D:2 * D].view(B, S, 8, 32).transpose(1, 2)
vh = qkvf[..., 2 * D:].view(B, S, 8, 32).transpose(1, 2)
of = ((qh @ kh.transpose(-1, -2)).softmax(-1) @ vh).transpose(1, 2).reshape(B, S, D)
of.retain_grad()
pf = xf + TO.lin(Pd, p + ".self_attn.out_proj", of)
pf.retain_grad()
yo = TO.ln(Pd, p + ".norm1", pf, 1e-5)
yo.backward(dy.float().view(B, S, D))
print("fwd qkv", rel(qkv.view(B, S, -1), qkvf), "o", rel(o.view(B, S, D), of), "pre", rel(pre.view(B, S, D), pf), "y", rel(y.view(B, S, D), yo))
print("bwd dpre", rel(dpre.view(B, S, D), pf.grad), "do", rel(do.view(B, S, D), of.grad), "dqkv", rel(dqkv.view(B, S, -1), qkvf.grad),
      "dq", rel(dqkv[:, :D].reshape(B, S, D), qkvf.grad[..., :D]), "dk", rel(dqkv[:, D:2 * D].reshape(B, S, D), qkvf.grad[..., D:2 * D]),
      "dv", rel(dqkv[:, 2 * D:].reshape(B, S, D), qkvf.grad[..., 2 * D:]))
print("dx", rel(dx2.view(B, S, D), xf.grad), "norms dpre", pf.grad.norm().item(), "dx", xf.grad.norm().item(),
      "dqkv", qkvf.grad.norm().item(), "dq", qkvf.grad[..., :D].norm().item(), "dk", qkvf.grad[..., D:2*D].norm().item(), "dv", qkvf.grad[..., 2*D:].norm().item())
print("scores scale: q std", qkvf[..., :D].std().item(), "k std", qkvf[..., D:2 * D].std().item())
