"""Row-tile-resident fused sub-layer kernels (csrc/layer_umma.cu) against plain torch fp32 arithmetic on identical
bf16-representable inputs.  Reference semantics: TransformerEncoderLayer.forward_post transformer.py:148-161,
TransformerDecoderLayer.forward_post transformer.py:211-232.

Tolerance: the kernels keep fp32 accumulators and round the hidden activation to bf16 once (it is an MMA operand);
outputs are stored in bf16 -> relative L2 <= 6e-3 (one bf16 rounding is 2^-9 = 2e-3 rms).  Train-mode dropout is
checked with the masks the kernels themselves regenerate, exported by gpvb200_dropout_mask."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _mlp_inputs(cuda, M, dff, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).to(cuda)
    x = r(M, 256).to(BF)
    w1 = (r(dff, 256) / 16).to(BF)
    w2 = (r(256, dff) / dff ** 0.5).to(BF)
    b1, b2 = 0.1 * r(dff), 0.1 * r(256)
    gamma, beta = 1 + 0.1 * r(256), 0.1 * r(256)
    return x, w1, b1, w2, b2, gamma, beta


def _mlp_ref(x, w1, b1, w2, b2, gamma, beta, eps, mask_h=None, mask_o=None, scale=1.0):
    h = torch.relu(x.float() @ w1.float().t() + b1)
    if mask_h is not None:
        h = h * mask_h * scale
    hb = h.to(BF).float()                       # the kernel feeds FFN2 with the bf16 hidden activation
    f = hb @ w2.float().t() + b2
    if mask_o is not None:
        f = f * mask_o * scale
    pre = x.float() + f
    mean = pre.mean(1, keepdim=True)
    var = pre.var(1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = (pre - mean) * rstd * gamma + beta
    return y, h, pre, torch.cat([mean, rstd], 1)


@pytest.mark.parametrize("M,dff,seq", [(9600, 2048, 0), (1200, 2048, 300), (3200, 2048, 100), (77, 128, 0), (128, 64, 0)])
def test_mlp_block_fwd(cuda, M, dff, seq):
    from gpv1_b200 import kernels as k
    x, w1, b1, w2, b2, gamma, beta = _mlp_inputs(cuda, M, dff, 1)
    y, h, pre, st = k.mlp_block_fwd(x, w1, b1, w2, b2, gamma, beta, 1e-5, seq_len=seq)
    torch.cuda.synchronize()
    ry, rh, rpre, rst = _mlp_ref(x, w1, b1, w2, b2, gamma, beta, 1e-5)
    errs = {"y": _rel(y, ry), "h": _rel(h, rh), "pre": _rel(pre, rpre), "mean": _rel(st[:, 0], rst[:, 0]), "rstd": _rel(st[:, 1], rst[:, 1])}
    print("mlp_block_fwd", (M, dff, seq), {kk: f"{v:.2e}" for kk, v in errs.items()})
    assert errs["y"] < 6e-3 and errs["h"] < 6e-3 and errs["pre"] < 6e-3, errs
    assert errs["mean"] < 2e-3 and errs["rstd"] < 2e-3, errs
    # inference form: nothing saved, same y
    y2, h2, pre2, st2 = k.mlp_block_fwd(x, w1, b1, w2, b2, gamma, beta, 1e-5, seq_len=seq, save=False)
    assert h2 is None and pre2 is None and st2 is None and torch.equal(y, y2)


def test_mlp_block_fwd_matches_unfused_kernels(cuda):
    """Same inputs through the three-launch path (GEMM + GEMM + LayerNorm) the engine used before."""
    from gpv1_b200 import kernels as k
    M, dff = 9600, 2048
    x, w1, b1, w2, b2, gamma, beta = _mlp_inputs(cuda, M, dff, 2)
    y, h, pre, st = k.mlp_block_fwd(x, w1, b1, w2, b2, gamma, beta, 1e-5)
    h0 = k.linear(x, w1, b1, act=k.ACT_RELU)
    pre0 = k.linear(h0, w2, b2, residual=x)
    y0, st0 = k.layernorm_fwd(pre0, gamma, beta, 1e-5)
    assert (h != h0).float().mean().item() < 1e-3 and _rel(h, h0) < 1e-3   # same k order, same rounding: (nearly) bit-identical
    assert _rel(pre, pre0) < 4e-3 and _rel(y, y0) < 6e-3 and _rel(st, st0) < 2e-3


def test_mlp_block_fwd_dropout(cuda):
    from gpv1_b200 import kernels as k
    M, dff = 1200, 2048
    x, w1, b1, w2, b2, gamma, beta = _mlp_inputs(cuda, M, dff, 3)
    seed = torch.tensor([5], dtype=torch.int64, device=cuda)
    dh, do = k.Drop(seed, 41, 0.1), k.Drop(seed, 42, 0.1)
    y, h, pre, st = k.mlp_block_fwd(x, w1, b1, w2, b2, gamma, beta, 1e-5, seq_len=300, drop_h=dh, drop_o=do)
    mh, mo = k.dropout_mask(M, dff, dh).float(), k.dropout_mask(M, 256, do).float()
    ry, rh, rpre, _ = _mlp_ref(x, w1, b1, w2, b2, gamma, beta, 1e-5, mh, mo, dh.scale)
    assert h.float()[mh == 0].abs().max().item() == 0
    assert _rel(h, rh) < 6e-3 and _rel(pre, rpre) < 6e-3 and _rel(y, ry) < 6e-3
    # the GEMM-epilogue dropout of the unfused path draws the same masks from the same (seed, site, row, col)
    h0 = k.linear(x, w1, b1, act=k.ACT_RELU, drop=dh, drop_mode=k.DROP_POST_ACT)
    assert (h != h0).float().mean().item() < 1e-3 and _rel(h, h0) < 1e-3


@pytest.mark.parametrize("M,seq,alpha", [(9600, 0, 1.0), (300, 0, 1.0 / 0.9), (640, 320, 1.0), (130, 0, 1.0)])
def test_mlp_block_bwd(cuda, M, seq, alpha):
    """dh = alpha (dy W2) (*) [h > 0], dx = dh W1 + dres in one launch, against the torch composition (dh rounded to bf16 in between, as
    the kernel hands it to the second contraction)."""
    from gpv1_b200 import kernels as k
    g = torch.Generator(device="cpu").manual_seed(M)
    r = lambda *s_: torch.randn(*s_, generator=g).to(cuda)
    dff = 2048
    dy, dres = (0.1 * r(M, 256)).to(BF), (0.1 * r(M, 256)).to(BF)
    w1, w2 = (r(dff, 256) / 16).to(BF), (r(256, dff) / 45).to(BF)
    h = torch.relu(r(M, dff)).to(BF)
    h[:, 5] = 0
    dx, dh = k.mlp_block_bwd(dy, w2.t().contiguous(), w1.t().contiguous(), h, dres, alpha=alpha, seq_len=seq)
    torch.cuda.synchronize()
    rdh = alpha * (dy.float() @ w2.float()) * (h.float() > 0)
    rdx = rdh.to(BF).float() @ w1.float() + dres.float()
    e_h, e_x = _rel(dh, rdh), _rel(dx, rdx)
    print("mlp_block_bwd", (M, seq), f"dh {e_h:.2e} dx {e_x:.2e}")
    assert e_h < 4e-3 and e_x < 4e-3, (e_h, e_x)
    assert dh.float()[h.float() == 0].abs().max().item() == 0


# ---------------------------------------------------------------------------------------------------- attention block
def _attn_inputs(cuda, B, Sq, Sk, seed, self_attn):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).to(cuda)
    if self_attn:
        qkv = r(B * Sq, 768).to(BF)
        q, k, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
    else:
        q = r(B * Sq, 256).to(BF)
        kv = r(B * Sk, 512).to(BF)
        k, v = kv[:, :256], kv[:, 256:]
    return q, k, v


def _attn_ref(q, k, v, B, Sq, Sk, scale, kmask=None, pmask=None, pscale=1.0):
    H, dh = 8, 32
    qh = q.float().view(B, Sq, H, dh).transpose(1, 2)
    kh = k.float().view(B, Sk, H, dh).transpose(1, 2)
    vh = v.float().view(B, Sk, H, dh).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    if kmask is not None:
        s = s.masked_fill(kmask.bool()[:, None, None, :], float("-inf"))
    lse2 = torch.logsumexp(s, -1) * 1.4426950408889634            # log2 domain, like the kernels
    pr = s.softmax(-1)
    if pmask is not None:
        pr = pr * pmask.view(B, H, Sq, Sk) * pscale
    o = (pr.to(BF).float() @ vh).transpose(1, 2).reshape(B * Sq, H * dh)   # the kernel feeds P V with bf16 probabilities
    return o, lse2


@pytest.mark.parametrize("B,Sq,Sk,self_attn,masked", [(4, 300, 300, True, False), (3, 100, 300, False, True), (5, 100, 100, True, False),
                                                      (2, 300, 300, True, True), (1, 37, 50, False, False), (2, 130, 16, False, False)])
def test_attn_block_core(cuda, B, Sq, Sk, self_attn, masked):
    from gpv1_b200 import kernels as k
    q, kk, v = _attn_inputs(cuda, B, Sq, Sk, 4, self_attn)
    kmask = None
    if masked:
        kmask = torch.zeros(B, Sk, dtype=torch.uint8, device=cuda)
        for b in range(B):
            kmask[b, Sk - 7 * (b + 1):] = 1
        kmask[0, 3] = 1
    scale = 32 ** -0.5
    o, lse = k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=scale, key_mask=kmask)
    torch.cuda.synchronize()
    ro, rlse = _attn_ref(q, kk, v, B, Sq, Sk, scale, kmask)
    e_o, e_l = _rel(o, ro), (lse - rlse).abs().max().item()
    o0, lse0 = k.attention_fwd(q, kk, v, B=B, H=8, Sq=Sq, Sk=Sk, dh=32, scale=scale, key_mask=kmask)     # the mma.sync kernel it replaces
    print("attn_block core", (B, Sq, Sk), f"o {e_o:.2e} lse {e_l:.2e} vs mma.sync kernel: o {_rel(o, o0):.2e} lse {(lse - lse0).abs().max().item():.2e}")
    assert e_o < 6e-3 and e_l < 2e-3, (e_o, e_l)
    assert _rel(o, o0) < 6e-3 and (lse - lse0).abs().max().item() < 2e-3


def test_attn_block_dropout_uses_the_backward_kernels_mask(cuda):
    """Train mode: the probabilities' dropout mask must be the one attention_bwd regenerates (row = (b H + h) Sq + q, col = key)."""
    from gpv1_b200 import kernels as k
    B, Sq, Sk = 3, 300, 300
    q, kk, v = _attn_inputs(cuda, B, Sq, Sk, 5, True)
    seed = torch.tensor([9], dtype=torch.int64, device=cuda)
    dp = k.Drop(seed, 77, 0.1)
    scale = 32 ** -0.5
    o, lse = k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=scale, drop_p=dp)
    pm = k.dropout_mask(B * 8 * Sq, Sk, dp).float()
    ro, rlse = _attn_ref(q, kk, v, B, Sq, Sk, scale, None, pm, dp.scale)
    o0, lse0 = k.attention_fwd(q, kk, v, B=B, H=8, Sq=Sq, Sk=Sk, dh=32, scale=scale, drop=dp)
    assert _rel(o, ro) < 8e-3 and (lse - rlse).abs().max().item() < 2e-3
    assert _rel(o, o0) < 8e-3


@pytest.mark.parametrize("B,Sq,Sk,self_attn", [(4, 300, 300, True), (3, 100, 300, False), (2, 100, 100, True)])
def test_attn_block_fused_tail(cuda, B, Sq, Sk, self_attn):
    """y = LN(x + o Wo^T + bo) in the same kernel, against the torch composition on the kernel's own o."""
    from gpv1_b200 import kernels as k
    q, kk, v = _attn_inputs(cuda, B, Sq, Sk, 6, self_attn)
    g = torch.Generator(device="cpu").manual_seed(7)
    r = lambda *s: torch.randn(*s, generator=g).to(cuda)
    wo = (r(256, 256) / 16).to(BF)
    bo, gamma, beta = 0.1 * r(256), 1 + 0.1 * r(256), 0.1 * r(256)
    x = r(B * Sq, 256).to(BF)
    scale = 32 ** -0.5
    y, o, lse, pre, st = k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=scale, wo=wo, bo=bo, x=x, gamma=gamma, beta=beta)
    torch.cuda.synchronize()
    ro, rlse = _attn_ref(q, kk, v, B, Sq, Sk, scale)
    assert _rel(o, ro) < 6e-3 and (lse - rlse).abs().max().item() < 2e-3
    rpre = x.float() + o.float() @ wo.float().t() + bo
    mean, var = rpre.mean(1, keepdim=True), rpre.var(1, unbiased=False, keepdim=True)
    ry = (rpre - mean) * (var + 1e-5).rsqrt() * gamma + beta
    errs = {"pre": _rel(pre, rpre), "y": _rel(y, ry), "mean": _rel(st[:, 0], mean[:, 0]), "rstd": _rel(st[:, 1], (var + 1e-5).rsqrt()[:, 0])}
    print("attn_block fused", (B, Sq, Sk), {kk_: f"{v_:.2e}" for kk_, v_ in errs.items()})
    assert errs["pre"] < 6e-3 and errs["y"] < 6e-3 and errs["mean"] < 2e-3 and errs["rstd"] < 2e-3, errs
    y2, o2, lse2, pre2, st2 = k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=scale, wo=wo, bo=bo, x=x, gamma=gamma, beta=beta, save=False)
    assert o2 is None and pre2 is None and torch.equal(y, y2)
