"""Decode-loop bookkeeping kernels (csrc/decode.cu) against the reference's algorithm (exp/gpv/models/gpv.py:178-196, 256-328) run
with torch ops on the SAME logits: integer outputs (token ids, parents, sequences) must be bit-exact, scores within fp32
round-off of log_softmax.  This isolates the bookkeeping from the bf16 noise of the decoder that feeds it."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_beam_step(lg, score, ids, t, K):
    """One iteration of the reference's loop (as restated in oracle/torch_oracle.py:beam_search) on given logits [B,K,V]."""
    B = lg.shape[0]
    top = torch.log_softmax(lg, -1).topk(K, -1)
    cand = score[:, :, None] + top.values
    if t == 0:
        cand[:, 1:] = cand[:, 1:] * 0 - 1e9
    flat = cand.reshape(B, K * K)
    order = torch.sort(flat, dim=1, descending=True, stable=True).indices[:, :K]
    k1 = order // K
    new_last = torch.gather(top.indices.reshape(B, K * K), 1, order)
    ids = torch.cat((torch.gather(ids, 1, k1[:, :, None].expand(-1, -1, ids.shape[2])), new_last[:, :, None]), 2)
    return ids, torch.gather(flat, 1, order), k1


@pytest.mark.parametrize("B,K,V,L,ld", [(4, 5, 512, 5, 512), (64, 5, 8192, 20, 8192), (3, 1, 97, 4, 104), (2, 8, 1000, 6, 1024)])
def test_beam_update_matches_reference_algorithm(cuda, B, K, V, L, ld):
    from gpv1_b200 import kernels as k
    g = torch.Generator().manual_seed(B * 131 + K)
    ids = [torch.full((B, K, L), 1, dtype=torch.int64, device=cuda) for _ in range(2)]
    score = [torch.zeros((B, K), device=cuda) for _ in range(2)]
    parent = torch.empty(B * K, dtype=torch.int64, device=cuda)
    tok = torch.empty(B * K, dtype=torch.int64, device=cuda)
    r_ids = torch.full((B, K, 1), 1, dtype=torch.int64)
    r_score = torch.zeros(B, K)
    worst = 0.0
    for t in range(L - 1):
        lg = torch.zeros(B * K, ld)
        lg[:, :V] = 3 * torch.randn(B * K, V, generator=g)
        lg[:, V:] = 1e30                                            # padding columns past V must never be read
        dl = lg.to(cuda)
        k.beam_update(dl, V, t, score[t & 1], ids[t & 1], score[(t + 1) & 1], ids[(t + 1) & 1], parent, tok)
        r_ids, r_score, k1 = _ref_beam_step(lg[:, :V].view(B, K, V), r_score, r_ids, t, K)
        got = ids[(t + 1) & 1].cpu()
        assert torch.equal(got[:, :, :t + 2], r_ids), t
        assert torch.equal(tok.cpu().view(B, K), r_ids[:, :, -1])
        assert torch.equal(parent.cpu().view(B, K), torch.arange(B)[:, None] * K + k1)
        err = (score[(t + 1) & 1].cpu() - r_score).abs().max().item()
        worst = max(worst, err)
        assert err <= 2e-5 * max(1.0, r_score.abs().max().item()), (t, err)
    print(f"[parity] beam_update B={B} K={K} V={V}: sequences / parents / tokens exact over {L - 1} steps, worst |d log-prob| {worst:.2e}")


def test_beam_update_tie_rule(cuda):
    """Ties: equal logits -> the lower token id first within a hypothesis (topk), equal candidates -> the first in (k1 major, k2 minor)
    order (stable sort): the reference's behaviour on exactly representable values."""
    from gpv1_b200 import kernels as k
    B, K, V, L = 2, 3, 64, 3
    lg = torch.zeros(B * K, V)
    lg[:, 7] = 2.0
    lg[:, 3] = 2.0                       # tie between tokens 3 and 7
    lg[:, 40] = 1.0
    ids = [torch.full((B, K, L), 1, dtype=torch.int64, device=cuda) for _ in range(2)]
    score = [torch.zeros((B, K), device=cuda) for _ in range(2)]
    parent = torch.empty(B * K, dtype=torch.int64, device=cuda)
    tok = torch.empty(B * K, dtype=torch.int64, device=cuda)
    k.beam_update(lg.to(cuda), V, 0, score[0], ids[0], score[1], ids[1], parent, tok)
    assert tok.cpu().view(B, K).tolist() == [[3, 7, 40]] * B
    assert parent.cpu().view(B, K).tolist() == [[0, 0, 0], [3, 3, 3]]
    # t = 1: every hypothesis has the same score and the same logits -> candidates tie across k1; the first K in flat order win
    score[1].fill_(0.5)
    k.beam_update(lg.to(cuda), V, 1, score[1], ids[1], score[0], ids[0], parent, tok)
    assert tok.cpu().view(B, K).tolist() == [[3, 7, 3]] * B
    assert parent.cpu().view(B, K).tolist() == [[0, 0, 1], [3, 3, 4]]
    assert ids[0].cpu()[0].tolist() == [[1, 3, 3], [1, 3, 7], [1, 7, 3]]


@pytest.mark.parametrize("rows,V,ld,masked", [(5, 512, 512, False), (64, 8192, 8192, True), (3, 97, 104, True)])
def test_argmax_with_vocab_mask(cuda, rows, V, ld, masked):
    from gpv1_b200 import kernels as k
    g = torch.Generator().manual_seed(rows)
    lg = torch.full((rows, ld), 1e30)
    lg[:, :V] = torch.randn(rows, V, generator=g)
    lg[0, 5] = lg[0, 9] = 50.0                                       # tie: first index wins
    vm = None
    if masked:
        vm = torch.full((V,), -10000.0)
        vm[torch.randperm(V, generator=g)[: max(4, V // 8)]] = 0
        vm[5] = vm[9] = 0
    out = torch.empty((rows, 3, V), device=cuda)
    ids = k.argmax(lg.to(cuda), V, vocab_mask=vm.to(cuda) if masked else None, out=out[:, 1])
    ref = lg[:, :V] + (vm if masked else 0)
    assert torch.equal(out[:, 1].cpu(), ref)
    assert torch.equal(ids.cpu(), ref.argmax(-1))
    assert ids[0].item() == 5


def test_reorder_rows(cuda):
    from gpv1_b200 import kernels as k
    rows, L, D, t = 20, 7, 768, 3
    src = torch.randn(rows, L, D, device=cuda).to(torch.bfloat16)
    dst = torch.full_like(src, 7.0)
    parent = torch.randint(0, rows, (rows,), device=cuda)
    k.reorder_rows(src, dst, parent, t * D)
    assert torch.equal(dst[:, :t], src.index_select(0, parent)[:, :t])
    assert (dst[:, t:] == 7.0).all()                                 # positions not decoded yet are not touched


@pytest.mark.parametrize("B,rep,H,Sk,dh,L", [(64, 5, 8, 120, 96, 0), (320, 1, 8, 7, 96, 20), (3, 2, 12, 33, 64, 0), (1, 1, 8, 1, 96, 20), (2, 3, 16, 20, 48, 0)])
def test_decode_attention(cuda, B, rep, H, Sk, dh, L):
    """One query row per hypothesis against cached K / V, `rep` hypotheses sharing a K / V batch (the beams of an image) or per-hypothesis
    caches read in place with a batch stride (L > 0: cache [Bq, L, 3 D], keys 0..Sk-1), against torch."""
    from gpv1_b200 import kernels as k
    torch.manual_seed(B * 7 + rep)
    D = H * dh
    Bq = B if rep == 1 else B * rep
    nb = Bq // rep
    q = torch.randn(Bq, D, device=cuda).to(torch.bfloat16)
    if L:
        cache = torch.randn(nb, L, 3 * D, device=cuda).to(torch.bfloat16)
        c2 = cache.view(nb * L, 3 * D)
        kk, vv, bs = c2[:, D:2 * D], c2[:, 2 * D:], L * 3 * D
        kf, vf = cache[:, :Sk, D:2 * D].float(), cache[:, :Sk, 2 * D:].float()
    else:
        kv = torch.randn(nb * Sk, 2 * D, device=cuda).to(torch.bfloat16)
        kk, vv, bs = kv[:, :D], kv[:, D:], Sk * 2 * D
        kf, vf = kv[:, :D].float().view(nb, Sk, D), kv[:, D:].float().view(nb, Sk, D)
    scale = dh ** -0.5
    o = k.decode_attention(q, kk, vv, Bq=Bq, rep=rep, H=H, Sk=Sk, dh=dh, scale=scale, bs_k=bs, bs_v=bs)
    qh = q.float().view(nb, rep, H, dh)
    kh, vh = kf.view(nb, Sk, H, dh), vf.view(nb, Sk, H, dh)
    sc = torch.einsum("brhd,bshd->brhs", qh, kh) * scale
    ref = torch.einsum("brhs,bshd->brhd", sc.softmax(-1), vh).reshape(Bq, D)
    err = ((o.float() - ref).norm() / ref.norm()).item()
    assert err < 5e-3, err
