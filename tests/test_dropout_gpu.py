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


@pytest.mark.parametrize("M,D", [(640, 768), (3203, 256), (9600, 256), (37, 128)])
def test_layernorm_dropout_outputs(cuda, M, D):
    from gpv1_b200 import kernels as k
    torch.manual_seed(3)
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


# ---------------------------------------------------------------------------------------------------- engine blocks
@pytest.fixture(scope="module")
def ctx(cuda):
    import json
    import os
    from gpv1_b200.config import load_config
    from gpv1_b200.model import GPV
    from oracle import torch_oracle as TO
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gpv_specs.json")))
    P = TO.make_state([tuple(s) for s in g["specs"]], seed=0)
    vocab = ["__pad__", "__cls__", "__stop__", "__unk__"] + [f"w{i}" for i in range(g["V"] - 4)]
    model = GPV(load_config().model, vocab=vocab, vocab_embed=P["answer_head.vocab_embed"].numpy())
    model.load_state_dict(P, strict=True)
    model.to(cuda)
    eng = model.engine
    eng.refresh()
    Pd = {n: (t.to(cuda).to(BF).float() if t.dtype.is_floating_point and t.dim() >= 2 else t.to(cuda)) for n, t in P.items()}
    return model, eng, Pd


@pytest.mark.parametrize("B,S", [(2, 64), (22, 300)])
def test_encoder_layer_with_dropout_matches_autograd(ctx, cuda, B, S):
    """One DETR encoder layer (transformer.py:148-161) in train mode: attention-probability dropout, dropout1, the FFN's
    hidden dropout and dropout2, forward and backward, against autograd fed with the masks of the engine's sites.  (22, 300) is
    large enough (66 row tiles) for the tcgen05 attn_block / mlp_block kernels to be the ones that run."""
    import torch.nn.functional as F
    from gpv1_b200 import kernels as k
    model, eng, Pd = ctx
    torch.manual_seed(0)
    D, H = 256, 8
    p = "detr.transformer.encoder.layers.1"
    x = (torch.randn(B * S, D, device=cuda)).to(BF)
    pos = torch.randn(S, D, device=cuda).to(BF)
    dy = (0.1 * torch.randn(B * S, D, device=cuda)).to(BF)
    eng.train_mode = True
    eng.drop_seed.fill_(123)
    try:
        eng.grad_arena.zero_()
        y1, sa = eng._self_attn_fwd(p, x, pos, S, B, S, H)
        y2, sf = eng._ffn_fwd(p + ".linear1", p + ".linear2", p + ".norm2", y1, 1e-5)
        d1 = eng._ffn_bwd(p + ".linear1", p + ".linear2", p + ".norm2", dy, sf)
        dx = eng._self_attn_bwd(p, d1, sa, None, B, S, H)
        eng._join()
        m = lambda site, rows, N: k.dropout_mask(rows, N, eng._drop(site)).float()
        mp = m(f"{p}.self_attn.probs", B * H * S, S).view(B, H, S, S)
        m1 = m(f"{p}.norm1.in", B * S, D).view(B, S, D)
        mh = m(f"{p}.linear1.hidden", B * S, 2048).view(B, S, 2048)
        m2 = m(f"{p}.norm2.in", B * S, D).view(B, S, D)
    finally:
        eng.train_mode = False
    sc = 1.0 / 0.9
    Pl = {n: t.clone().requires_grad_(True) for n, t in Pd.items() if n.startswith(p) and t.dtype.is_floating_point}
    # each sub-block gets the engine's own bf16 input (as tests/test_blocks_gpu.py does: ReLU gates flip under input noise)
    xf = x.float().view(B, S, D).requires_grad_(True)
    qk = xf + pos.float()[None]
    Wi, bi = Pl[f"{p}.self_attn.in_proj_weight"], Pl[f"{p}.self_attn.in_proj_bias"]
    q = F.linear(qk, Wi[:D], bi[:D]).view(B, S, H, D // H).transpose(1, 2) * (D // H) ** -0.5
    kk = F.linear(qk, Wi[D:2 * D], bi[D:2 * D]).view(B, S, H, D // H).transpose(1, 2)
    v = F.linear(xf, Wi[2 * D:], bi[2 * D:]).view(B, S, H, D // H).transpose(1, 2)
    a = ((torch.softmax(q @ kk.transpose(-1, -2), -1) * mp * sc) @ v).transpose(1, 2).reshape(B, S, D)
    a = F.linear(a, Pl[f"{p}.self_attn.out_proj.weight"], Pl[f"{p}.self_attn.out_proj.bias"])
    h1 = F.layer_norm(xf + a * m1 * sc, (D,), Pl[f"{p}.norm1.weight"], Pl[f"{p}.norm1.bias"], 1e-5)
    h1.backward(d1.float().view(B, S, D))                      # upstream gradient = the engine's own FFN-block output
    y1f = y1.float().view(B, S, D).requires_grad_(True)
    f = F.relu(F.linear(y1f, Pl[f"{p}.linear1.weight"], Pl[f"{p}.linear1.bias"])) * mh * sc
    f = F.linear(f, Pl[f"{p}.linear2.weight"], Pl[f"{p}.linear2.bias"])
    out = F.layer_norm(y1f + f * m2 * sc, (D,), Pl[f"{p}.norm2.weight"], Pl[f"{p}.norm2.bias"], 1e-5)
    out.backward(dy.float().view(B, S, D))
    TOL = 1.5e-2
    assert _rel(y1.view(B, S, D), h1) < TOL
    assert _rel(y2.view(B, S, D), out) < TOL
    assert _rel(d1.view(B, S, D), y1f.grad) < TOL
    assert _rel(dx.view(B, S, D), xf.grad) < TOL
    bad = []
    for n, t in Pl.items():
        if t.grad is not None and n in eng.G and t.grad.norm() > 1e-6:
            r = _rel(eng.G[n], t.grad)
            if r > TOL:
                bad.append((n, r))
    assert not bad, bad


def test_train_mode_step_runs_and_eval_is_unchanged(ctx, cuda):
    """model.train(): a full step with every dropout site active gives a finite loss that changes from step to step
    (new masks) and finite gradients; model.eval() afterwards reproduces the dropout-free loss bit for bit."""
    from oracle.make_golden import make_inputs
    model, eng, Pd = ctx
    images, qids, ans, targets = make_inputs(3, 192, 256, 8, 5, 55, ["CocoCaptioning", "CocoVqa", "CocoDetection"])
    dt = [{kk: (v.to(cuda) if torch.is_tensor(v) else v) for kk, v in t.items()} for t in targets]
    b = (images.to(cuda), qids.to(cuda), ans.to(cuda), dt)

    def step():
        for p_ in model.parameters():
            p_.grad = None
        loss = model(*b)
        loss.backward()
        torch.cuda.synchronize()
        gn = sum(p_.grad.float().norm().item() ** 2 for p_ in model.parameters() if p_.grad is not None) ** 0.5
        return loss.item(), gn

    model.eval()
    l_eval, g_eval = step()
    model.train()
    try:
        l1, g1 = step()
        l2, g2 = step()
    finally:
        model.eval()
    l_eval2, _ = step()
    for v_ in (l1, l2, g1, g2):
        assert v_ == v_ and abs(v_) < 1e6
    assert l1 != l2 and l1 != l_eval
    assert abs(l1 - l_eval) < 0.5 * abs(l_eval) and abs(l2 - l_eval) < 0.5 * abs(l_eval)
    assert 0.2 * g_eval < g1 < 5 * g_eval
    assert abs(l_eval2 - l_eval) <= 1e-6 * abs(l_eval)
