"""Attention / LayerNorm / helper / criterion kernels vs. plain PyTorch fp32 references of the same ops.

Tolerances (bf16 storage, fp32 accumulate): 2e-2 of the reference's max magnitude for bf16 outputs, 2e-3 for fp32 ones.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=2e-2, name=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{name}: max err {err:.4g} vs scale {scale:.4g}"


def _bf(*shape, dev, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("B,H,Sq,Sk,dh,causal,masked", [
    (2, 8, 300, 300, 32, False, True), (2, 8, 100, 100, 32, False, False), (2, 8, 100, 300, 32, False, True),
    (3, 16, 20, 100, 48, False, False), (3, 16, 100, 20, 48, False, False), (2, 8, 20, 20, 96, True, False),
    (2, 8, 19, 120, 96, False, False), (2, 12, 20, 20, 64, False, True), (1, 8, 7, 7, 96, True, False)])
def test_attention(cuda, B, H, Sq, Sk, dh, causal, masked):
    from gpv1_b200 import kernels as k
    torch.manual_seed(0)
    D = H * dh
    # packed projections, like the model produces them: q | k | v side by side when Sq == Sk
    q = _bf(B * Sq, D, dev=cuda)
    kv = _bf(B * Sk, 2 * D, dev=cuda)
    kk, vv = kv[:, :D], kv[:, D:]
    km = None
    if masked:
        km = torch.zeros(B, Sk, dtype=torch.uint8, device=cuda)
        km[0, Sk - 3:] = 1
        km[-1, : Sk // 4] = 1
    scale = dh ** -0.5
    o, lse = k.attention_fwd(q, kk, vv, B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=scale, causal=causal, key_mask=km)

    def ref(qf, kf, vf):
        qh = qf.view(B, Sq, H, dh).transpose(1, 2)
        kh = kf.view(B, Sk, H, dh).transpose(1, 2)
        vh = vf.view(B, Sk, H, dh).transpose(1, 2)
        s = (qh @ kh.transpose(-1, -2)) * scale
        if km is not None:
            s = s.masked_fill(km.bool()[:, None, None, :], float("-inf"))
        if causal:
            s = s.masked_fill(torch.ones(Sq, Sk, device=cuda).triu(1).bool(), float("-inf"))
        return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B * Sq, D)

    qf = q.float().requires_grad_(True)
    kf = kk.float().requires_grad_(True)
    vf = vv.float().requires_grad_(True)
    r = ref(qf, kf, vf)
    _close(o, r, name="attn fwd")
    d_o = _bf(B * Sq, D, dev=cuda)
    r.backward(d_o.float())
    dq = torch.empty(B * Sq, D, device=cuda, dtype=torch.bfloat16)
    dkv = torch.empty(B * Sk, 2 * D, device=cuda, dtype=torch.bfloat16)
    k.attention_bwd(q, kk, vv, o, d_o, lse, dq, dkv[:, :D], dkv[:, D:], B=B, H=H, Sq=Sq, Sk=Sk, dh=dh, scale=scale,
                    causal=causal, key_mask=km)
    _close(dq, qf.grad, tol=3e-2, name="attn dq")
    _close(dkv[:, :D], kf.grad, tol=3e-2, name="attn dk")
    _close(dkv[:, D:], vf.grad, tol=3e-2, name="attn dv")


@pytest.mark.parametrize("M,D,eps,affine", [(9600, 256, 1e-5, True), (640, 768, 1e-12, True), (3200, 2048, 1e-5, False),
                                            (77, 768, 1e-5, True), (5, 256, 1e-5, False), (3203, 256, 1e-5, True), (50, 64, 1e-5, True)])
def test_layernorm(cuda, M, D, eps, affine):
    from gpv1_b200 import kernels as k
    torch.manual_seed(1)
    x = _bf(M, D, dev=cuda, scale=3.0)
    g = (torch.rand(D, device=cuda) + 0.5) if affine else None
    b = torch.randn(D, device=cuda) if affine else None
    y, stats = k.layernorm_fwd(x, g, b, eps)
    xf = x.float().requires_grad_(True)
    gf = g.clone().requires_grad_(True) if affine else None
    bf = b.clone().requires_grad_(True) if affine else None
    r = F.layer_norm(xf, (D,), gf, bf, eps)
    _close(y, r, name="ln fwd")
    dy = _bf(M, D, dev=cuda)
    r.backward(dy.float())
    dg = torch.zeros(D, device=cuda) if affine else None
    db = torch.zeros(D, device=cuda) if affine else None
    dx = k.layernorm_bwd(dy, x, stats, g, dg, db)
    _close(dx, xf.grad, name="ln dx")
    if affine:
        _close(dg, gf.grad, tol=5e-3, name="ln dgamma")
        _close(db, bf.grad, tol=5e-3, name="ln dbeta")
    # strided output into a wider buffer (ROI features next to hs, detr_roi_head.py:92)
    wide = torch.zeros(M, D + 256, device=cuda, dtype=torch.bfloat16)
    k.layernorm_fwd(x, g, b, eps, out=wide[:, :D], need_stats=False)
    _close(wide[:, :D], r, name="ln strided")
    assert wide[:, D:].abs().max().item() == 0


def test_helpers(cuda):
    from gpv1_b200 import kernels as k
    torch.manual_seed(2)
    B, S, D = 3, 300, 256
    x, p = _bf(B * S, D, dev=cuda), _bf(S, D, dev=cuda)
    _close(k.add_rowbcast(x, p), x.float() + p.float().repeat(B, 1), name="add_rowbcast")
    _close(k.add_rowbcast(None, p, M=B * S), p.float().repeat(B, 1), name="rowbcast")
    dy = _bf(1000, 776, dev=cuda)
    out = torch.ones(776, device=cuda)
    k.colsum(dy, out)
    _close(out, 1 + dy.float().sum(0), tol=2e-3, name="colsum")
    out = torch.zeros(S, D, device=cuda)
    k.batch_reduce(x, out, B, S)
    _close(out, x.float().view(B, S, D).sum(0), tol=2e-3, name="batch_reduce")
    img = _bf(2, 30, 41, 64, dev=cuda)
    mp = k.maxpool3x3s2(img)
    ref = F.max_pool2d(img.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.float(), ref)
    im = torch.randn(2, 3, 37, 50, device=cuda)
    col, Ho, Wo = k.stem_im2col(im)
    w = torch.randn(64, 3, 7, 7, device=cuda)
    wp = w.permute(0, 2, 3, 1).reshape(64, 147)
    got = (col[:, :147].float() @ wp.t()).view(2, Ho, Wo, 64)
    ref = F.conv2d(im.to(torch.bfloat16).float(), w, stride=2, padding=3).permute(0, 2, 3, 1)
    _close(got, ref, tol=2e-3, name="stem im2col")
    assert col[:, 147:].abs().max().item() == 0
    tab = torch.randn(50, 768, device=cuda)
    ids = torch.randint(0, 50, (40,), device=cuda)
    pos = torch.randn(20, 768, device=cuda)
    cst = torch.randn(768, device=cuda)
    _close(k.gather_rows(tab, ids, pos=pos, cst=cst, T=20), tab[ids] + pos.repeat(2, 1) + cst, name="gather")
    src = _bf(2 * 100, 768, dev=cuda)
    mem = torch.zeros(2 * 120, 768, device=cuda, dtype=torch.bfloat16)
    k.copy_rows(src, mem, 200, 768, dst_map=(100, 120, 0))
    assert torch.equal(mem.view(2, 120, 768)[:, :100], src.view(2, 100, 768))
    back = torch.zeros_like(src)
    k.copy_rows(mem, back, 200, 768, src_map=(100, 120, 0))
    assert torch.equal(back, src)
    f = torch.randn(1001, device=cuda)
    assert torch.equal(k.cast_bf16(f), f.to(torch.bfloat16))


def test_roi_weights_vs_torchvision(cuda):
    """mean over the 7x7 aligned ROI-align bins == Wroi @ features (detr_roi_head.py:44-56)."""
    import torchvision
    from gpv1_b200 import kernels as k
    torch.manual_seed(3)
    B, Q, C, H, W = 2, 100, 64, 15, 20
    feat = torch.randn(B, C, H, W, device=cuda)
    boxes = torch.cat([torch.rand(B, Q, 2, device=cuda), 0.02 + 0.9 * torch.rand(B, Q, 2, device=cuda)], -1)
    sb = torch.zeros_like(boxes)
    sb[:, :, 0] = W * (boxes[:, :, 0] - 0.5 * boxes[:, :, 2])
    sb[:, :, 1] = H * (boxes[:, :, 1] - 0.5 * boxes[:, :, 3])
    sb[:, :, 2] = W * (boxes[:, :, 0] + 0.5 * boxes[:, :, 2])
    sb[:, :, 3] = H * (boxes[:, :, 1] + 0.5 * boxes[:, :, 3])
    ref = torchvision.ops.roi_align(feat, list(torch.unbind(sb)), output_size=7, aligned=True)
    ref = ref.view(B, Q, C, 7, 7).mean(-1).mean(-1)
    wr = k.roi_weights(boxes.view(B * Q, 4).contiguous(), H, W, 304).float().view(B, Q, 304)[:, :, :H * W]
    got = torch.bmm(wr, feat.permute(0, 2, 3, 1).reshape(B, H * W, C))
    _close(got, ref, tol=1e-2, name="roi weights")  # weights are bf16-rounded


def test_relevance_mix(cuda):
    from gpv1_b200 import kernels as k
    torch.manual_seed(4)
    B, Q, T, D = 3, 100, 20, 768
    x = _bf(B * Q, D, dev=cuda)
    logits = torch.zeros(B * Q, 8, device=cuda)
    logits[:, :2] = torch.randn(B * Q, 2, device=cuda)
    tok = 0.1 * torch.randn(2, D, device=cuda)
    mem = torch.zeros(B * (Q + T), D, device=cuda, dtype=torch.bfloat16)
    k.relevance_mix_fwd(x, logits, tok, mem, G=Q, out_gstride=Q + T, out_off=0)
    lf = logits[:, :2].clone().requires_grad_(True)
    tf = tok.clone().requires_grad_(True)
    ref = x.float() + lf.softmax(-1) @ tf
    _close(mem.view(B, Q + T, D)[:, :Q].reshape(B * Q, D), ref, name="mix fwd")
    dmem = _bf(B * (Q + T), D, dev=cuda)
    ref.backward(dmem.view(B, Q + T, D)[:, :Q].reshape(B * Q, D).float())
    dl = torch.zeros(B * Q, 8, device=cuda)
    dt = torch.zeros(2, D, device=cuda)
    k.relevance_mix_bwd(dmem, logits, tok, dl, dt, M=B * Q, G=Q, gstride=Q + T, off=0)
    _close(dl[:, :2], lf.grad, tol=5e-3, name="mix dlogits")
    _close(dt, tf.grad, tol=5e-3, name="mix dtok")


def test_cross_entropy(cuda):
    from gpv1_b200 import kernels as k
    torch.manual_seed(5)
    rows, V = 57, 1000
    logits = 3 * torch.randn(rows, V, device=cuda)
    tg = torch.randint(0, V, (rows,), device=cuda)
    w = torch.rand(rows, device=cuda)
    w[::5] = 0
    loss = torch.zeros(1, device=cuda)
    dl = torch.empty(rows, V, device=cuda, dtype=torch.bfloat16)
    rl = torch.empty(rows, device=cuda)
    k.ce_fwd_bwd(logits, tg, w, loss, dl, rl)
    lf = logits.clone().requires_grad_(True)
    per = F.cross_entropy(lf, tg, reduction="none")
    ref = (per * w).sum()
    ref.backward()
    _close(rl, per, tol=1e-4, name="ce rows")
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    _close(dl, lf.grad, name="ce dlogits")


def test_set_criterion(cuda):
    """Matched-pair losses and gradients vs. autograd on the reference formulas (set_criterion.py:44-97)."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(6)
    B, Q = 4, 100
    tc = [5, 0, 3, 7]
    loc_valid = torch.tensor([1, 1, 0, 1], dtype=torch.uint8, device=cuda)
    off = torch.tensor([0, 5, 5, 8, 15], dtype=torch.int32, device=cuda)
    tb = torch.cat([0.2 + 0.6 * torch.rand(15, 2, device=cuda), 0.05 + 0.3 * torch.rand(15, 2, device=cuda)], -1)
    logits = torch.zeros(B * Q, 8, device=cuda)
    logits[:, :2] = torch.randn(B * Q, 2, device=cuda)
    pre = torch.randn(B * Q, 4, device=cuda)
    boxes = torch.zeros(B * Q, 8, device=cuda)
    boxes[:, :4] = pre.sigmoid()
    Kmax = 7
    iq = torch.full((B, Kmax), -1, dtype=torch.int64, device=cuda)
    it = torch.full((B, Kmax), -1, dtype=torch.int64, device=cuda)
    for b in range(B):
        qs = torch.randperm(Q, device=cuda)[: tc[b]].sort().values
        iq[b, : tc[b]] = qs
        it[b, : tc[b]] = torch.randperm(tc[b], device=cuda)
    nvalid_img = [0, 1, 3]
    n_match = sum(tc[b] for b in nvalid_img)
    wsum = n_match * 1.0 + (len(nvalid_img) * Q - n_match) * 0.1
    num_boxes = float(max(n_match, 1))
    out3 = torch.zeros(3, device=cuda)
    dl = torch.zeros(B * Q, 8, device=cuda)
    dbp = torch.zeros(B * Q, 8, device=cuda, dtype=torch.bfloat16)
    k.set_criterion(logits, boxes, tb, off, iq, it, loc_valid, eos_coef=0.1, weight_sum=wsum, num_boxes=num_boxes, wt_ce=1.0,
                    wt_bbox=5.0, wt_giou=2.0, out3=out3, dlogits=dl, dbox_pre=dbp)

    # reference formulas
    lf = logits[:, :2].clone().view(B, Q, 2).requires_grad_(True)
    pf = pre.clone().view(B, Q, 4).requires_grad_(True)
    bx = pf.sigmoid()
    sel = nvalid_img
    tgt_cls = torch.ones(B, Q, dtype=torch.int64, device=cuda)
    src, tgt = [], []
    for b in sel:
        tgt_cls[b, iq[b, : tc[b]]] = 0
        src.append(bx[b, iq[b, : tc[b]]])
        tgt.append(tb[off[b]: off[b + 1]][it[b, : tc[b]]])
    src, tgt = torch.cat(src), torch.cat(tgt)
    l_ce = F.cross_entropy(lf[sel].transpose(1, 2), tgt_cls[sel], torch.tensor([1.0, 0.1], device=cuda))
    l_l1 = F.l1_loss(src, tgt, reduction="none").sum() / num_boxes

    def xyxy(b):
        return torch.stack([b[:, 0] - 0.5 * b[:, 2], b[:, 1] - 0.5 * b[:, 3], b[:, 0] + 0.5 * b[:, 2], b[:, 1] + 0.5 * b[:, 3]], -1)

    a, c = xyxy(src), xyxy(tgt)
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (c[:, 2] - c[:, 0]) * (c[:, 3] - c[:, 1])
    wh = (torch.min(a[:, 2:], c[:, 2:]) - torch.max(a[:, :2], c[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area1 + area2 - inter
    ewh = (torch.max(a[:, 2:], c[:, 2:]) - torch.min(a[:, :2], c[:, :2])).clamp(min=0)
    earea = ewh[:, 0] * ewh[:, 1]
    giou = inter / union - (earea - union) / earea
    l_gi = (1 - giou).sum() / num_boxes
    (1.0 * l_ce + 5.0 * l_l1 + 2.0 * l_gi).backward()
    ref3 = torch.stack([l_ce, l_l1, l_gi]).detach()
    _close(out3, ref3, tol=1e-4, name="set losses")
    _close(dl[:, :2], lf.grad.view(B * Q, 2), tol=1e-3, name="set dlogits")
    _close(dbp[:, :4], pf.grad.view(B * Q, 4), tol=1e-2, name="set dbox")


@pytest.mark.parametrize("H,W", [(37, 50), (64, 96), (33, 47)])
def test_stem_s2d_conv_matches_conv2d(cuda, H, W):
    """7x7/s2/p3 stem (backbone.py:72 -> torchvision conv1) as space-to-depth + 4-tap K=64 implicit GEMM over
    overlapping TMA rows, against F.conv2d on the same bf16-rounded operands (even, odd and ragged sizes)."""
    import torch.nn.functional as F
    from gpv1_b200 import kernels as k
    torch.manual_seed(3)
    im = torch.randn(2, 3, H, W, device=cuda)
    w = torch.randn(64, 3, 7, 7, device=cuda) * 0.1
    b = torch.randn(64, device=cuda)
    xv, Ho, Wo = k.stem_s2d(im)
    ws = k.stem_weight_s2d(w)
    y = k.conv(xv, ws, ksize=7, taps=k.STEM_TAPS, Ho=Ho, Wo=Wo, N=64, K=64, bias=b, act=k.ACT_RELU)
    ref = F.relu(F.conv2d(im.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b, stride=2, padding=3)).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, err


def test_clip_adamw_matches_torch(cuda):
    """optim.ClipAdamW (two launches over a flat gradient arena) against torch.nn.utils.clip_grad_norm_ +
    torch.optim.AdamW with the reference's groups (exp/gpv/train_distr.py:228-253, 414-428) over three steps."""
    from gpv1_b200.optim import ClipAdamW
    torch.manual_seed(5)
    shapes = {"detr.backbone.0.body.layer2.0.conv1.weight": (128, 256, 1, 1), "detr.transformer.encoder.layers.0.linear1.bias": (2048,),
              "detr_joiner.weight": (768, 2304), "bert_joiner.bias": (768,), "relevance_tokens": (2, 768), "odd.tensor": (5, 7, 3)}
    total = sum((math.prod(s) + 7) // 8 * 8 for s in shapes.values())
    arena = torch.zeros(total + 3, device=cuda)
    named, ref_params, off = [], {}, 0
    for n, s in shapes.items():
        num = math.prod(s)
        p = torch.randn(s, device=cuda)
        g = arena[off:off + num].view(s)
        off += (num + 7) // 8 * 8
        if n == "odd.tensor":
            off += 3                                            # an unaligned gradient offset: scalar path
        named.append((n, p, g))
        ref_params[n] = torch.nn.Parameter(p.clone())
    named[-1] = ("odd.tensor", named[-1][1], arena[total - 105 + 3 - 8:total - 8 + 3].view(5, 7, 3)) if False else named[-1]
    opt = ClipAdamW(named, arena, lr=1e-3, lr_backbone=1e-4, weight_decay=1e-2, clip_max_norm=0.1)
    groups = [[], [], [], []]
    from gpv1_b200.optim import group_of
    for n in shapes:
        groups[group_of(n)].append(ref_params[n])
    ref = torch.optim.AdamW([{"params": groups[0], "lr": 1e-4}, {"params": groups[1]}, {"params": groups[2]}, {"params": groups[3]}],
                            lr=1e-3, weight_decay=1e-2)
    for step in range(3):
        for n, p, g in named:
            g.copy_(torch.randn_like(g) * (0.01 if step == 1 else 1.0))   # step 1: norm below the threshold -> no clipping
            ref_params[n].grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(groups[0] + groups[1], 0.1)
        ref.step()
        opt.step()
        torch.cuda.synchronize()
        assert abs(opt.grad_norm().item() - norm.item()) <= 1e-4 * norm.item()
        for n, p, g in named:
            assert torch.allclose(g, ref_params[n].grad, rtol=1e-5, atol=1e-8), (step, n, "clipped grad")
            assert torch.allclose(p, ref_params[n].data, rtol=2e-5, atol=2e-6), (step, n, (p - ref_params[n].data).abs().max().item())


def test_clip_adamw_state_dict_is_torch_layout_and_steps_are_per_parameter(cuda):
    """(a) state_dict() is torch.optim.AdamW's own layout -- a torch AdamW built the reference's way
    (train_distr.py:234-253) loads it and continues identically, and ClipAdamW loads torch's; (b) a tensor whose state is missing
    from the loaded checkpoint (DETR parameters unfrozen for the second training phase) starts its bias correction at step 1,
    as torch does, instead of inheriting the global step."""
    from gpv1_b200.optim import ClipAdamW, group_of
    torch.manual_seed(6)
    shapes = {"detr.backbone.0.body.layer2.0.conv1.weight": (16, 8, 1, 1), "detr.class_embed.bias": (24,), "bert_joiner.weight": (8, 8),
              "relevance_tokens": (2, 8), "frozen.never.steps": (3,)}
    live = [n for n in shapes if n != "frozen.never.steps"]

    def build(names):
        arena = torch.zeros(sum(math.prod(shapes[n]) for n in names) + 8, device=cuda)
        named, off = [], 0
        for n in names:
            num = math.prod(shapes[n])
            named.append((n, torch.randn(shapes[n], device=cuda), arena[off:off + num].view(shapes[n])))
            off += num
        return ClipAdamW(named, arena, lr=1e-3, lr_backbone=1e-4, weight_decay=1e-2, clip_max_norm=0.1, all_names=list(shapes)), named

    def torch_opt(params):
        groups = [[], [], [], []]
        for n in shapes:
            groups[group_of(n)].append(params[n])
        return torch.optim.AdamW([{"params": groups[0], "lr": 1e-4}, {"params": groups[1]}, {"params": groups[2]}, {"params": groups[3]}],
                                 lr=1e-3, weight_decay=1e-2), groups

    def step_both(opt, named, ref, groups, params):
        for n, p, g in named:
            g.copy_(torch.randn_like(g))
            params[n].grad = g.clone()
        torch.nn.utils.clip_grad_norm_(groups[0] + groups[1], 0.1)
        ref.step()
        opt.step()

    # ---- phase 1: the detr.* tensors are frozen (no optimizer state), 4 steps
    phase1 = [n for n in live if not n.startswith("detr.")]
    opt1, named1 = build(phase1)
    params = {n: torch.nn.Parameter(torch.randn(shapes[n], device=cuda)) for n in shapes}
    for n, p, g in named1:
        params[n].data.copy_(p)
    ref, groups = torch_opt(params)
    for _ in range(4):
        step_both(opt1, named1, ref, groups, params)
    sd = opt1.state_dict()
    ref_sd = ref.state_dict()
    assert set(sd["state"]) == set(ref_sd["state"])                    # same indices present (only the tensors that stepped)
    assert [g["params"] for g in sd["param_groups"]] == [g["params"] for g in ref_sd["param_groups"]]
    for i, st in ref_sd["state"].items():
        assert float(sd["state"][i]["step"]) == float(st["step"]) == 4.0
        assert torch.allclose(sd["state"][i]["exp_avg"], st["exp_avg"], rtol=1e-5, atol=1e-8)
    ref2, groups2 = torch_opt(params)
    ref2.load_state_dict({"state": sd["state"], "param_groups": sd["param_groups"]})      # torch reads ClipAdamW's file
    # ---- phase 2: everything trains; ClipAdamW resumes from TORCH's state dict
    opt2, named2 = build(live)
    for n, p, g in named2:
        p.copy_(params[n].data)
    opt2.load_state_dict(ref_sd)
    assert opt2.t == 4 and [opt2.t - s0 for s0 in opt2.step0] == [0 if n.startswith("detr.") else 4 for n in live]
    for _ in range(3):
        step_both(opt2, named2, ref2, groups2, params)
        torch.cuda.synchronize()
        for n, p, g in named2:
            assert torch.allclose(p, params[n].data, rtol=2e-5, atol=2e-6), (n, (p - params[n].data).abs().max().item())
    sd2 = opt2.state_dict()
    assert float(sd2["state"][opt2.index_of["detr.class_embed.bias"]]["step"]) == 3.0
    assert float(sd2["state"][opt2.index_of["relevance_tokens"]]["step"]) == 7.0


def test_stem_s2d_uint8_equals_normalised_fp32(cuda):
    """uint8 NHWC ingest: (u8/255 - mean)/std folded into the stem's read gives the same s2d map as the
    fp32 NCHW path fed with the reference's ToTensor + Normalize output (coco_generic_dataset.py:31-32)."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(9)
    u8 = torch.randint(0, 256, (2, 37, 50, 3), device=cuda, dtype=torch.uint8)
    mean = torch.tensor(k.IMAGENET_MEAN, device=cuda).view(1, 3, 1, 1)
    std = torch.tensor(k.IMAGENET_STD, device=cuda).view(1, 3, 1, 1)
    f32 = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    a, Ho, Wo = k.stem_s2d(u8)
    b, Ho2, Wo2 = k.stem_s2d(f32.contiguous())
    assert (Ho, Wo) == (Ho2, Wo2)
    # compare the underlying [B, Hp, Wp, 16] maps (the 64-wide views overlap)
    ma, mb = a[..., :16].float(), b[..., :16].float()
    assert (ma - mb).abs().max().item() <= 2e-2        # one bf16 ulp at |x| <= 2.7 where the two fp32 roundings differ
    assert (ma != mb).float().mean().item() < 0.05
