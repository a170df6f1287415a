"""Thin tensor-level wrappers over the C-ABI (no autograd here; see ops.py).

Every function takes CUDA torch tensors, passes raw pointers + the current stream to libgpvb200.so and returns
torch tensors allocated by PyTorch's caching allocator.  Nothing here computes on the host.
"""
import ctypes

import torch

from . import _C

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3
AUX_NONE, AUX_RELU_MASK, AUX_GELU_GRAD = 0, 1, 2
BF16 = torch.bfloat16


class Drop:
    """One dropout site of a training step: (device uint64 step counter, site id, probability).  The mask is a
    counter-based function of these and of the element's (row, col) -- csrc/common.cuh -- so backward regenerates it."""
    __slots__ = ("seed", "site", "p")

    def __init__(self, seed, site, p):
        self.seed, self.site, self.p = seed, int(site) & 0xFFFFFFFF, float(p)

    @property
    def scale(self):
        return 1.0 / (1.0 - self.p)


DROP_PRE_RESIDUAL, DROP_POST_ACT = 1, 2


def dropout_mask(rows, N, drop):
    """uint8 [rows, N]: 1 = kept (test / debug export of the mask every kernel regenerates)."""
    out = torch.empty((rows, N), device=drop.seed.device, dtype=torch.uint8)
    _C.check(_C.lib().gpvb200_dropout_mask(_C.ptr(out), ctypes.c_int64(rows), N, _C.ptr(drop.seed), ctypes.c_uint32(drop.site),
                                           ctypes.c_float(drop.p), _C.stream_ptr()), "dropout_mask")
    return out


def _req(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("gpvb200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return t


LANE_CTA_CAP = 0      # > 0 while the engine issues work on a lane (Engine._aside): cap of the persistent grid of atomically accumulating launches


def _launch_gemm(d: _C.GemmDesc):
    if LANE_CTA_CAP > 0 and d.d_atomic:
        d.max_ctas = LANE_CTA_CAP
    _C.check(_C.lib().gpvb200_gemm(ctypes.byref(d), _C.stream_ptr()), "gemm")


def gemm(A, B, D, *, M, N, K, lda, ldb, ldd, a_mn=False, b_mn=False, batch=1, a_bs=0, b_bs=0, d_bs=0,
         bias=None, rowscale=None, residual=None, ldr=0, aux=None, ldaux=0, aux_mode=AUX_NONE, act=ACT_NONE,
         alpha=1.0, atomic=False, splits=1, D2=None, drop=None, drop_mode=0):
    """mode-0 contraction (see include/gpvb200.h). D dtype decides bf16 / fp32 output."""
    d = _C.GemmDesc()
    d.mode = 0
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.act, d.aux_mode = act, aux_mode
    d.d_fp32 = int(D.dtype == torch.float32)
    d.d_atomic = int(atomic)
    d.splits = splits
    d.alpha = alpha
    d.A, d.B, d.D, d.D2 = _C.ptr(A), _C.ptr(B), _C.ptr(D), _C.ptr(D2)
    d.bias, d.rowscale = _C.ptr(_req(bias, torch.float32)), _C.ptr(_req(rowscale, torch.float32))
    d.residual, d.aux = _C.ptr(_req(residual)), _C.ptr(_req(aux, BF16))
    d.res_fp32 = int(residual is not None and residual.dtype == torch.float32)
    d.lda, d.ldb, d.ldd, d.ldr, d.ldaux = lda, ldb, ldd, ldr, ldaux
    d.a_batch_stride, d.b_batch_stride, d.d_batch_stride = a_bs, b_bs, d_bs
    if drop is not None and drop_mode:
        d.drop_seed, d.drop_mode, d.drop_site, d.drop_p = _C.ptr(drop.seed), drop_mode, drop.site, drop.p
    _launch_gemm(d)
    return D


def linear(x, w, bias=None, *, act=ACT_NONE, residual=None, out_dtype=BF16, out=None, out2=None, alpha=1.0, drop=None, drop_mode=0):
    """y[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias + residual).  x, w bf16 row-major (last dim contiguous)."""
    _req(x, BF16), _req(w, BF16)
    M, K = x.shape
    N = w.shape[0]
    y = out if out is not None else torch.empty((M, N), device=x.device, dtype=out_dtype)
    return gemm(x, w, y, M=M, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldd=y.stride(0), bias=bias,
                residual=residual, ldr=residual.stride(0) if residual is not None else 0, act=act, D2=out2, alpha=alpha,
                drop=drop, drop_mode=drop_mode)


def linear_dgrad(dy, w, *, aux=None, aux_mode=AUX_NONE, residual=None, out=None, out_dtype=BF16, alpha=1.0):
    """dx[M,K] = (dy[M,N] @ w[N,K] + residual) (* mask(aux)).  w is read in its forward layout (MN-major B)."""
    _req(dy, BF16), _req(w, BF16)
    M, N = dy.shape
    K = w.shape[1]
    dx = out if out is not None else torch.empty((M, K), device=dy.device, dtype=out_dtype)
    return gemm(dy, w, dx, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldd=dx.stride(0), b_mn=True,
                residual=residual, ldr=residual.stride(0) if residual is not None else 0,
                aux=aux, ldaux=aux.stride(0) if aux is not None else 0, aux_mode=aux_mode, alpha=alpha)


def linear_wgrad(dy, x, dw, *, rowscale=None, splits=0):
    """dw[N,K] (fp32) += dy[M,N]^T @ x[M,K]; both operands read in place (MN-major), split over M (splits=0: the
    library picks tile width and split count from its cost model)."""
    _req(dy, BF16), _req(x, BF16), _req(dw, torch.float32)
    M, N = dy.shape
    K = x.shape[1]
    return gemm(dy, x, dw, M=N, N=K, K=M, lda=dy.stride(0), ldb=x.stride(0), ldd=dw.stride(0), a_mn=True, b_mn=True,
                rowscale=rowscale, atomic=True, splits=splits)


_TAPS3 = [(r - 1, s - 1) for r in range(3) for s in range(3)]


def conv(x, w, *, ksize, stride=1, bias=None, act=ACT_NONE, residual=None, aux=None, aux_mode=AUX_NONE, out=None,
         taps=None, tap_w=None, b_mn=False, out_geom=None, Ho=None, Wo=None, N=None, K=None):
    """NHWC implicit-GEMM convolution (mode 1).

    x [n,Hi,Wi,C] bf16; w [taps,N,K] bf16 (b_mn=False) or [taps,K,N] (b_mn=True, data-gradient form).
    taps: list of (dh, dw) input offsets relative to ho*stride, wo*stride (default: 1x1 -> [(0,0)], 3x3 pad 1).
    out_geom: (OH, OW, out_stride, off_h, off_w) to scatter the tile into a larger output (stride-2 dgrad).
    """
    _req(x, BF16), _req(w, BF16)
    n, Hi, Wi, C = x.shape
    if taps is None:
        taps = [(0, 0)] if ksize == 1 else _TAPS3
    if tap_w is None:
        tap_w = list(range(len(taps)))
    if Ho is None:
        Ho = (Hi - 1) // stride + 1
        Wo = (Wi - 1) // stride + 1
    if N is None:
        N = w.shape[2] if b_mn else w.shape[1]
    if K is None:
        K = C
    d = _C.GemmDesc()
    d.mode = 1
    d.N, d.K = N, K
    d.batch = 1
    d.b_mn = int(b_mn)
    d.act, d.aux_mode = act, aux_mode
    d.n_img, d.Hi, d.Wi, d.Ho, d.Wo, d.stride = n, Hi, Wi, Ho, Wo, stride
    d.ntaps = len(taps)
    for i, (dh, dw) in enumerate(taps):
        d.tap_dh[i], d.tap_dw[i], d.tap_w[i] = dh, dw, tap_w[i]
    if out_geom is None:
        OH, OW, os_, oh, ow = Ho, Wo, 1, 0, 0
    else:
        OH, OW, os_, oh, ow = out_geom
    d.OH, d.OW, d.out_stride, d.out_off_h, d.out_off_w = OH, OW, os_, oh, ow
    y = out if out is not None else torch.empty((n, OH, OW, N), device=x.device, dtype=BF16)
    d.d_fp32 = int(y.dtype == torch.float32)
    d.alpha = 1.0
    d.A, d.B, d.D = _C.ptr(x), _C.ptr(w), _C.ptr(y)
    d.bias = _C.ptr(_req(bias, torch.float32))
    d.residual, d.aux = _C.ptr(_req(residual, BF16)), _C.ptr(_req(aux, BF16))
    d.lda, d.ldb, d.ldd = x.stride(2), w.stride(1), y.stride(2)
    d.ldr = residual.stride(2) if residual is not None else 0
    d.ldaux = aux.stride(2) if aux is not None else 0
    d.b_batch_stride = w.stride(0)
    _launch_gemm(d)
    return y


def conv_wgrad(dy, x, dw, *, ksize, stride=1, rowscale=None, splits=0):
    """dw[taps,Cout,Cin] (fp32) += sum_pixels dy[n,ho,wo,:]^T x[n,ho*s+dh,wo*s+dw,:]  (mode 2)."""
    _req(dy, BF16), _req(x, BF16), _req(dw, torch.float32)
    n, Ho, Wo, Cout = dy.shape
    _, Hi, Wi, Cin = x.shape
    taps = [(0, 0)] if ksize == 1 else _TAPS3
    d = _C.GemmDesc()
    d.mode = 2
    d.M, d.N = Cout, Cin
    d.batch = 1
    d.d_fp32, d.d_atomic = 1, 1
    d.n_img, d.Hi, d.Wi, d.Ho, d.Wo, d.stride = n, Hi, Wi, Ho, Wo, stride
    d.ntaps = len(taps)
    for i, (dh, dw_) in enumerate(taps):
        d.tap_dh[i], d.tap_dw[i], d.tap_w[i] = dh, dw_, i
    if splits <= 0:
        tiles = ((Cout + 127) // 128) * ((Cin + 127) // 128) * len(taps)
        splits = max(1, (2 * 148 + tiles - 1) // tiles)
    d.splits = splits
    d.alpha = 1.0
    d.A, d.B, d.D = _C.ptr(dy), _C.ptr(x), _C.ptr(dw)
    d.rowscale = _C.ptr(_req(rowscale, torch.float32))
    d.lda, d.ldb, d.ldd = dy.stride(2), x.stride(2), dw.stride(1)
    d.d_batch_stride = dw.stride(0)
    _launch_gemm(d)
    return dw


# ------------------------------------------------------------------------------------------------ matcher
def matcher_cost(logits, boxes, tgt_boxes, tgt_labels, tgt_offsets, Tmax, w_class, w_bbox, w_giou, *, C=None):
    """fp32 cost blocks [B,Q,Tmax] (utils/matcher.py:53-72).  logits/boxes are [B,Q,>=C] / [B,Q,>=4] with a contiguous
    last dimension (padded rows are read in place through their row stride)."""
    B, Q = logits.shape[:2]
    C = C if C is not None else logits.shape[2]
    cost = torch.zeros((B, Q, max(Tmax, 1)), device=logits.device, dtype=torch.float32)
    if B and Tmax:
        _C.check(_C.lib().gpvb200_matcher_cost(
            _C.ptr(_req(logits, torch.float32)), ctypes.c_int64(logits.stride(1)), _C.ptr(_req(boxes, torch.float32)),
            ctypes.c_int64(boxes.stride(1)), _C.ptr(_req(tgt_boxes, torch.float32)),
            _C.ptr(_req(tgt_labels, torch.int64)), _C.ptr(_req(tgt_offsets, torch.int32)), B, Q, C, Tmax,
            ctypes.c_float(w_class), ctypes.c_float(w_bbox), ctypes.c_float(w_giou), _C.ptr(cost), _C.stream_ptr()),
            "matcher_cost")
    return cost[:, :, :Tmax] if Tmax else cost[:, :, :0]


def lsap(cost, tgt_offsets):
    """cost [B,Q,Tmax] fp32 contiguous -> (idx_q, idx_t) int64 [B, min(Q,Tmax)] on device, -1 padded."""
    B, Q, Tmax = cost.shape
    K = min(Q, Tmax)
    oq = torch.full((B, K), -1, device=cost.device, dtype=torch.int64)
    ot = torch.full((B, K), -1, device=cost.device, dtype=torch.int64)
    if B and K:
        cost = cost.contiguous()
        _C.check(_C.lib().gpvb200_lsap(_C.ptr(_req(cost, torch.float32)), _C.ptr(_req(tgt_offsets, torch.int32)), B, Q, Tmax,
                                       _C.ptr(oq), _C.ptr(ot), _C.stream_ptr()), "lsap")
    return oq, ot


# ------------------------------------------------------------------------------------------------ attention
def attention_fwd(q, k, v, *, B, H, Sq, Sk, dh, scale, causal=False, key_mask=None, need_lse=True, out=None, bs_k=0, bs_v=0,
                  drop=None):
    """q/k/v: bf16 2-D views [B*S, >=H*dh] (row stride = tokens' leading dimension; may be slices of a packed QKV).
    bs_k / bs_v: batch strides in elements when K / V are read in place from a cache [B][S_max][...] with Sk <= S_max.
    Returns (o [B*Sq, H*dh] bf16, lse [B,H,Sq] fp32 log2-domain or None)."""
    o = out if out is not None else torch.empty((B * Sq, H * dh), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, Sq), device=q.device, dtype=torch.float32) if need_lse else None
    i64 = ctypes.c_int64
    if drop is not None:
        assert not bs_k and not bs_v
        _C.check(_C.lib().gpvb200_attention_fwd_drop(
            _C.ptr(_req(q, BF16)), _C.ptr(_req(k, BF16)), _C.ptr(_req(v, BF16)), _C.ptr(o), _C.ptr(lse),
            _C.ptr(_req(key_mask, torch.uint8)), i64(q.stride(0)), i64(k.stride(0)), i64(v.stride(0)), i64(o.stride(0)),
            B, H, Sq, Sk, dh, int(causal), ctypes.c_float(scale), _C.ptr(drop.seed), ctypes.c_uint32(drop.site),
            ctypes.c_float(drop.p), _C.stream_ptr()), "attention_fwd_drop")
        return o, lse
    _C.check(_C.lib().gpvb200_attention_fwd_bs(
        _C.ptr(_req(q, BF16)), _C.ptr(_req(k, BF16)), _C.ptr(_req(v, BF16)), _C.ptr(o), _C.ptr(lse),
        _C.ptr(_req(key_mask, torch.uint8)), i64(q.stride(0)), i64(k.stride(0)), i64(v.stride(0)), i64(o.stride(0)),
        i64(0), i64(bs_k), i64(bs_v), i64(0), B, H, Sq, Sk, dh, int(causal), ctypes.c_float(scale), _C.stream_ptr()), "attention_fwd")
    return o, lse


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, *, B, H, Sq, Sk, dh, scale, causal=False, key_mask=None, drop=None):
    """Writes dq/dk/dv (bf16 2-D views, may be slices of one packed gradient buffer)."""
    i64 = ctypes.c_int64
    if drop is not None:
        _C.check(_C.lib().gpvb200_attention_bwd_drop(
            _C.ptr(_req(q, BF16)), _C.ptr(_req(k, BF16)), _C.ptr(_req(v, BF16)), _C.ptr(_req(o, BF16)), _C.ptr(_req(d_o, BF16)),
            _C.ptr(_req(lse, torch.float32)), _C.ptr(_req(key_mask, torch.uint8)), _C.ptr(dq), _C.ptr(dk), _C.ptr(dv),
            i64(q.stride(0)), i64(k.stride(0)), i64(v.stride(0)), i64(o.stride(0)), i64(d_o.stride(0)), i64(dq.stride(0)),
            i64(dk.stride(0)), i64(dv.stride(0)), B, H, Sq, Sk, dh, int(causal), ctypes.c_float(scale), _C.ptr(drop.seed),
            ctypes.c_uint32(drop.site), ctypes.c_float(drop.p), _C.stream_ptr()), "attention_bwd_drop")
        return
    _C.check(_C.lib().gpvb200_attention_bwd(
        _C.ptr(_req(q, BF16)), _C.ptr(_req(k, BF16)), _C.ptr(_req(v, BF16)), _C.ptr(_req(o, BF16)), _C.ptr(_req(d_o, BF16)),
        _C.ptr(_req(lse, torch.float32)), _C.ptr(_req(key_mask, torch.uint8)), _C.ptr(dq), _C.ptr(dk), _C.ptr(dv),
        i64(q.stride(0)), i64(k.stride(0)), i64(v.stride(0)), i64(o.stride(0)), i64(d_o.stride(0)), i64(dq.stride(0)),
        i64(dk.stride(0)), i64(dv.stride(0)), B, H, Sq, Sk, dh, int(causal), ctypes.c_float(scale), _C.stream_ptr()),
        "attention_bwd")


# ------------------------------------------------------------------------------------------------ layernorm
def layernorm_fwd(x, gamma, beta, eps, *, out=None, need_stats=True, drop=None):
    """drop: y = dropout(LN(x)) (BERT embeddings)."""
    M, D = x.shape
    y = out if out is not None else torch.empty((M, D), device=x.device, dtype=BF16)
    stats = torch.empty((M, 2), device=x.device, dtype=torch.float32) if need_stats else None
    if drop is not None:
        _C.check(_C.lib().gpvb200_layernorm_fwd_drop(
            _C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0)), _C.ptr(gamma), _C.ptr(beta), ctypes.c_float(eps), _C.ptr(y),
            ctypes.c_int64(y.stride(0)), _C.ptr(stats), M, D, _C.ptr(drop.seed), ctypes.c_uint32(drop.site), ctypes.c_float(drop.p),
            _C.stream_ptr()), "layernorm_fwd_drop")
        return y, stats
    _C.check(_C.lib().gpvb200_layernorm_fwd(_C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0)), _C.ptr(gamma), _C.ptr(beta),
                                            ctypes.c_float(eps), _C.ptr(y), ctypes.c_int64(y.stride(0)), _C.ptr(stats), M, D,
                                            _C.stream_ptr()), "layernorm_fwd")
    return y, stats


def layernorm_bwd(dy, x, stats, gamma, dgamma, dbeta, *, out=None, drop=None):
    """drop: also returns dx (*) mask / (1 - p) -- the gradient of the dropped-out sub-layer output in
    y = LN(res + dropout(f)) -- as a second tensor: (dx, dx_masked)."""
    M, D = x.shape
    dx = out if out is not None else torch.empty((M, D), device=x.device, dtype=BF16)
    if drop is not None:
        dxm = torch.empty((M, D), device=x.device, dtype=BF16)
        _C.check(_C.lib().gpvb200_layernorm_bwd_drop(
            _C.ptr(_req(dy, BF16)), ctypes.c_int64(dy.stride(0)), _C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0)), _C.ptr(stats),
            _C.ptr(gamma), _C.ptr(dx), ctypes.c_int64(dx.stride(0)), _C.ptr(dgamma), _C.ptr(dbeta), M, D, _C.ptr(dxm),
            ctypes.c_int64(dxm.stride(0)), _C.ptr(drop.seed), ctypes.c_uint32(drop.site), ctypes.c_float(drop.p), _C.stream_ptr()),
            "layernorm_bwd_drop")
        return dx, dxm
    _C.check(_C.lib().gpvb200_layernorm_bwd(_C.ptr(_req(dy, BF16)), ctypes.c_int64(dy.stride(0)), _C.ptr(_req(x, BF16)),
                                            ctypes.c_int64(x.stride(0)), _C.ptr(stats), _C.ptr(gamma), _C.ptr(dx),
                                            ctypes.c_int64(dx.stride(0)), _C.ptr(dgamma), _C.ptr(dbeta), M, D, _C.stream_ptr()),
             "layernorm_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ fused sub-layers
def mlp_block_fwd(x, w1, b1, w2, b2, gamma, beta, eps, *, save=True, seq_len=0, drop_h=None, drop_o=None):
    """y = LN(x + drop_o(W2 drop_h(relu(W1 x + b1)) + b2)) in ONE kernel (d_model = 256).  Returns (y, h, pre, stats);
    h / pre / stats (what the backward pass reads) are None when save is False."""
    M, D = x.shape
    dff = w1.shape[0]
    dev = x.device
    y = torch.empty((M, D), device=dev, dtype=BF16)
    h = torch.empty((M, dff), device=dev, dtype=BF16) if save else None
    pre = torch.empty((M, D), device=dev, dtype=BF16) if save else None
    stats = torch.empty((M, 2), device=dev, dtype=torch.float32) if save else None
    d0 = drop_h if drop_h is not None else drop_o
    if drop_h is not None and drop_o is not None:
        assert drop_h.seed.data_ptr() == drop_o.seed.data_ptr()
    i64, f32, u32 = ctypes.c_int64, ctypes.c_float, ctypes.c_uint32
    _C.check(_C.lib().gpvb200_mlp_block_fwd(
        _C.ptr(_req(x, BF16)), i64(x.stride(0)), _C.ptr(_req(w1, BF16)), i64(w1.stride(0)), _C.ptr(_req(b1, torch.float32)),
        _C.ptr(_req(w2, BF16)), i64(w2.stride(0)), _C.ptr(_req(b2, torch.float32)), _C.ptr(_req(gamma, torch.float32)),
        _C.ptr(_req(beta, torch.float32)), f32(eps), _C.ptr(y), i64(y.stride(0)), _C.ptr(h), i64(dff), _C.ptr(pre), i64(D),
        _C.ptr(stats), i64(M), D, dff, seq_len, _C.ptr(d0.seed) if d0 is not None else ctypes.c_void_p(0),
        u32(drop_h.site if drop_h is not None else 0), f32(drop_h.p if drop_h is not None else 0.0),
        u32(drop_o.site if drop_o is not None else 0), f32(drop_o.p if drop_o is not None else 0.0), _C.stream_ptr()),
        "mlp_block_fwd")
    return y, h, pre, stats


def mlp_block_bwd(dy, w2t, w1t, h, dres, *, alpha=1.0, seq_len=0):
    """(dx, dh) of the fused FFN sub-layer: dh = alpha * (dy @ W2) * [h > 0], dx = dh @ W1 + dres.  w2t = W2^T [d_ff, 256], w1t = W1^T
    [256, d_ff] (bf16, packed transposed copies)."""
    M, D = dy.shape
    dff = w2t.shape[0]
    dev = dy.device
    dh = torch.empty((M, dff), device=dev, dtype=BF16)
    dx = torch.empty((M, D), device=dev, dtype=BF16)
    i64 = ctypes.c_int64
    _C.check(_C.lib().gpvb200_mlp_block_bwd(_C.ptr(_req(dy, BF16)), i64(dy.stride(0)), _C.ptr(_req(w2t, BF16)), i64(w2t.stride(0)),
                                            _C.ptr(_req(w1t, BF16)), i64(w1t.stride(0)), _C.ptr(_req(h, BF16)), i64(h.stride(0)),
                                            ctypes.c_float(alpha), _C.ptr(_req(dres, BF16)), i64(dres.stride(0)), _C.ptr(dh), i64(dff),
                                            _C.ptr(dx), i64(D), i64(M), D, dff, seq_len, _C.stream_ptr()), "mlp_block_bwd")
    return dx, dh


def attn_block_fwd(q, k, v, *, B, Sq, Sk, scale, key_mask=None, wo=None, bo=None, x=None, gamma=None, beta=None, eps=1e-5,
                   save=True, drop_p=None, drop_o=None, H=8, dh=32):
    """tcgen05 attention for d_model = 256 (8 x 32).  wo None: returns (o, lse).  Otherwise the output projection, residual and
    LayerNorm run in the same kernel: returns (y, o, lse, pre, stats) (o / lse / pre / stats None when save is False)."""
    dev = q.device
    D = H * dh
    fuse = wo is not None
    need = save or not fuse
    o = torch.empty((B * Sq, D), device=dev, dtype=BF16) if need else None
    lse = torch.empty((B, H, Sq), device=dev, dtype=torch.float32) if need else None
    y = torch.empty((B * Sq, D), device=dev, dtype=BF16) if fuse else None
    pre = torch.empty((B * Sq, D), device=dev, dtype=BF16) if fuse and save else None
    stats = torch.empty((B * Sq, 2), device=dev, dtype=torch.float32) if fuse and save else None
    d0 = drop_p if drop_p is not None else drop_o
    i64, f32, u32 = ctypes.c_int64, ctypes.c_float, ctypes.c_uint32
    _C.check(_C.lib().gpvb200_attn_block_fwd(
        _C.ptr(_req(q, BF16)), i64(q.stride(0)), _C.ptr(_req(k, BF16)), i64(k.stride(0)), _C.ptr(_req(v, BF16)), i64(v.stride(0)),
        _C.ptr(_req(key_mask, torch.uint8)), B, H, Sq, Sk, dh, f32(scale), _C.ptr(o), i64(D), _C.ptr(lse),
        _C.ptr(_req(wo, BF16)), i64(wo.stride(0) if fuse else 0), _C.ptr(_req(bo, torch.float32)), _C.ptr(_req(x, BF16)),
        i64(x.stride(0) if x is not None else 0), _C.ptr(_req(gamma, torch.float32)), _C.ptr(_req(beta, torch.float32)), f32(eps),
        _C.ptr(pre), i64(D), _C.ptr(y), i64(D), _C.ptr(stats), _C.ptr(d0.seed) if d0 is not None else ctypes.c_void_p(0),
        u32(drop_p.site if drop_p is not None else 0), f32(drop_p.p if drop_p is not None else 0.0),
        u32(drop_o.site if drop_o is not None else 0), f32(drop_o.p if drop_o is not None else 0.0), _C.stream_ptr()), "attn_block_fwd")
    if not fuse:
        return o, lse
    return y, o, lse, pre, stats


# ------------------------------------------------------------------------------------------------ decode bookkeeping
def argmax(logits, V, *, vocab_mask=None, out=None):
    """ids[row] = first arg-max of logits[row, :V] + vocab_mask; the masked logits are also written to out [rows, V] (any row stride)."""
    rows = logits.shape[0]
    ids = torch.empty((rows,), device=logits.device, dtype=torch.int64)
    _C.check(_C.lib().gpvb200_argmax(_C.ptr(_req(logits, torch.float32)), ctypes.c_int64(logits.stride(0)), rows, V,
                                     _C.ptr(_req(vocab_mask, torch.float32)), _C.ptr(_req(out, torch.float32)),
                                     ctypes.c_int64(out.stride(0) if out is not None else 0), _C.ptr(ids), _C.stream_ptr()), "argmax")
    return ids


def beam_update(logits, V, t, score_in, ids_in, score_out, ids_out, parent, tok, workspace=None):
    """One beam-search step (see gpvb200_beam_update): ids_* [B, K, L] int64, score_* [B, K] fp32, parent / tok [B*K] int64;
    workspace: int64 tensor of >= B*K*K elements (allocated here when None)."""
    B, K, L = ids_in.shape
    if workspace is None:
        workspace = torch.empty((B * K * K,), device=logits.device, dtype=torch.int64)
    assert workspace.numel() * 8 >= 8 * B * K * K
    _C.check(_C.lib().gpvb200_beam_update(_C.ptr(_req(logits, torch.float32)), ctypes.c_int64(logits.stride(0)), B, K, V, t, L,
                                          _C.ptr(_req(score_in, torch.float32)), _C.ptr(_req(ids_in, torch.int64)),
                                          _C.ptr(_req(score_out, torch.float32)), _C.ptr(_req(ids_out, torch.int64)),
                                          _C.ptr(_req(parent, torch.int64)), _C.ptr(_req(tok, torch.int64)), _C.ptr(_req(workspace, torch.int64)),
                                          _C.stream_ptr()), "beam_update")


def decode_attention(q, k, v, *, Bq, rep, H, Sk, dh, scale, bs_k, bs_v, out=None):
    """One query row per hypothesis against cached K / V (see gpvb200_decode_attention).  q [Bq, >= H*dh]; k / v: 2-D views whose row
    stride is the key stride, bs_k / bs_v the batch strides (elements); rep hypotheses share one K / V batch."""
    o = out if out is not None else torch.empty((Bq, H * dh), device=q.device, dtype=BF16)
    i64 = ctypes.c_int64
    _C.check(_C.lib().gpvb200_decode_attention(_C.ptr(_req(q, BF16)), i64(q.stride(0)), _C.ptr(_req(k, BF16)), i64(k.stride(0)), i64(bs_k),
                                               _C.ptr(_req(v, BF16)), i64(v.stride(0)), i64(bs_v), _C.ptr(o), i64(o.stride(0)), Bq, rep, H, Sk, dh,
                                               ctypes.c_float(scale), _C.stream_ptr()), "decode_attention")
    return o


def reorder_rows(src, dst, parent, n_elems):
    """dst[r, :n_elems] = src[parent[r], :n_elems] over the flattened rows of two [rows, ...] bf16 buffers."""
    rows = src.shape[0]
    row_elems = src.numel() // rows
    _C.check(_C.lib().gpvb200_reorder_rows(_C.ptr(_req(src, BF16)), _C.ptr(_req(dst, BF16)), _C.ptr(_req(parent, torch.int64)), rows,
                                           ctypes.c_int64(row_elems), ctypes.c_int64(n_elems), _C.stream_ptr()), "reorder_rows")


# ------------------------------------------------------------------------------------------------ helpers
def add_rowbcast(x, p, *, M=None, out=None):
    """out[m] = x[m] + p[m % P]; x may be None (pure broadcast of p over M rows)."""
    P, D = p.shape
    if x is not None:
        M = x.shape[0]
    y = out if out is not None else torch.empty((M, D), device=p.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_add_rowbcast(_C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0) if x is not None else 0),
                                           _C.ptr(_req(p, BF16)), ctypes.c_int64(p.stride(0)), _C.ptr(y),
                                           ctypes.c_int64(y.stride(0)), ctypes.c_int64(M), D, P, _C.stream_ptr()), "add_rowbcast")
    return y


def colsum(dy, out, N=None):
    """out[n] (fp32) += sum_m dy[m][n]."""
    M = dy.shape[0]
    N = N if N is not None else dy.shape[1]
    _C.check(_C.lib().gpvb200_colsum(_C.ptr(_req(dy, BF16)), ctypes.c_int64(dy.stride(0)), _C.ptr(_req(out, torch.float32)),
                                     ctypes.c_int64(M), N, _C.stream_ptr()), "colsum")


def batch_reduce(x, out, B, S):
    D = x.shape[1]
    _C.check(_C.lib().gpvb200_batch_reduce(_C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0)), _C.ptr(_req(out, torch.float32)),
                                           B, S, D, _C.stream_ptr()), "batch_reduce")


def maxpool3x3s2(x):
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), device=x.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_maxpool3x3s2(_C.ptr(_req(x, BF16)), _C.ptr(y), B, H, W, C, _C.stream_ptr()), "maxpool")
    return y


def stem_im2col(img, out=None):
    """img NCHW fp32 -> [B*Ho*Wo, 152] bf16 (7x7 stride 2 pad 3, k = tap*3 + c)."""
    B, C, H, W = img.shape
    assert C == 3
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    col = out if out is not None else torch.empty((B * Ho * Wo, 152), device=img.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_stem_im2col(_C.ptr(_req(img.contiguous(), torch.float32)), _C.ptr(col), B, H, W, _C.stream_ptr()),
             "stem_im2col")
    return col, Ho, Wo


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # coco_generic_dataset.py:32


def stem_s2d(img, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """img NCHW fp32 (already normalised) or NHWC uint8 (raw pixels; normalised here) -> (virtual NHWC view
    [B, Ho+4, Wo+4, 64] bf16 with pixel stride 16, Ho, Wo): the space-to-depth map of the 7x7/s2 stem; pixel (I, J) of
    the view = the 4 adjacent s2d pixels J..J+3 of row I (see elementwise.cu)."""
    if img.dtype == torch.uint8:
        B, H, W, C = img.shape
    else:
        B, C, H, W = img.shape
    assert C == 3
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Hp, Wp = Ho + 4, Wo + 4
    buf = torch.empty(B * Hp * Wp * 16 + 64, device=img.device, dtype=BF16)   # + 64: view pixels J > Wo (never stored) read up to 48 past the map
    if img.dtype == torch.uint8:
        f3 = ctypes.c_float * 3
        _C.check(_C.lib().gpvb200_stem_s2d_u8(_C.ptr(_req(img.contiguous())), _C.ptr(buf), B, H, W, f3(*mean), f3(*std), _C.stream_ptr()),
                 "stem_s2d_u8")
    else:
        _C.check(_C.lib().gpvb200_stem_s2d(_C.ptr(_req(img.contiguous(), torch.float32)), _C.ptr(buf), B, H, W, _C.stream_ptr()),
                 "stem_s2d")
    view = torch.as_strided(buf, (B, Hp, Wp, 64), (Hp * Wp * 16, Wp * 16, 16, 1))
    return view, Ho, Wo


STEM_TAPS = [(0, 0), (1, 0), (2, 0), (3, 0)]


def stem_weight_s2d(w_fold):
    """[64,3,7,7] (BN-folded, any float dtype) -> [4 taps a][64][64 = b*16 + dy*6 + dx*3 + c] bf16 for the s2d stem:
    r = 2a + dy - 1, s = 2b + dx - 1, zero where r or s falls outside 0..6 and in channels 12..15."""
    O = w_fold.shape[0]
    out = torch.zeros((4, O, 4, 16), device=w_fold.device, dtype=torch.float32)
    for a in range(4):
        for dy in range(2):
            r = 2 * a + dy - 1
            if not 0 <= r <= 6:
                continue
            for b in range(4):
                for dx in range(2):
                    s = 2 * b + dx - 1
                    if not 0 <= s <= 6:
                        continue
                    out[a, :, b, dy * 6 + dx * 3:dy * 6 + dx * 3 + 3] = w_fold[:, :, r, s].float()
    return out.reshape(4, O, 64).to(BF16).contiguous()


def roi_weights(boxes, H, W, ldw):
    """boxes [BQ, >=4] fp32 (cx,cy,w,h normalised) -> [BQ, ldw] bf16 separable ROI-mean weights."""
    BQ = boxes.shape[0]
    w = torch.empty((BQ, ldw), device=boxes.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_roi_weights(_C.ptr(_req(boxes, torch.float32)), ctypes.c_int64(boxes.stride(0)), _C.ptr(w),
                                          ctypes.c_int64(ldw), BQ, H, W, _C.stream_ptr()), "roi_weights")
    return w


def relevance_mix_fwd(x, logits, tok, out, *, G, out_gstride, out_off):
    M, D = x.shape
    _C.check(_C.lib().gpvb200_relevance_mix_fwd(_C.ptr(_req(x, BF16)), ctypes.c_int64(x.stride(0)), _C.ptr(_req(logits, torch.float32)),
                                                ctypes.c_int64(logits.stride(0)), _C.ptr(_req(tok, torch.float32)), _C.ptr(out),
                                                ctypes.c_int64(out.stride(0)), M, D, G, out_gstride, out_off, _C.stream_ptr()),
             "relevance_mix_fwd")


def relevance_mix_bwd(dy, logits, tok, dlogits, dtok, *, M, G, gstride, off):
    D = tok.shape[1]
    _C.check(_C.lib().gpvb200_relevance_mix_bwd(_C.ptr(_req(dy, BF16)), ctypes.c_int64(dy.stride(0)), _C.ptr(_req(logits, torch.float32)),
                                                ctypes.c_int64(logits.stride(0)), _C.ptr(_req(tok, torch.float32)),
                                                _C.ptr(_req(dlogits, torch.float32)), ctypes.c_int64(dlogits.stride(0)),
                                                _C.ptr(_req(dtok, torch.float32)), M, D, G, gstride, off, _C.stream_ptr()),
             "relevance_mix_bwd")


def gather_rows(table, ids, *, pos=None, cst=None, T=1, out=None, pad_mask=None, pad_id=0):
    """pad_mask: optional uint8 [M] output, 1 where ids == pad_id (the key-padding mask of the gathered token rows)."""
    M = ids.numel()
    D = table.shape[1]
    y = out if out is not None else torch.empty((M, D), device=table.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_gather_rows_mask(_C.ptr(_req(table, torch.float32)), _C.ptr(_req(ids, torch.int64)), _C.ptr(pos), _C.ptr(cst),
                                               _C.ptr(y), ctypes.c_int64(y.stride(0)), ctypes.c_int64(M), D, T,
                                               _C.ptr(_req(pad_mask, torch.uint8)), ctypes.c_int64(pad_id), _C.stream_ptr()),
             "gather_rows")
    return y


def copy_rows(src, dst, M, D, *, src_map=None, dst_map=None):
    """map = (G, gstride, off): row(m) = (m // G) * gstride + off + m % G; None = identity."""
    sG, sgs, so = src_map if src_map else (1 << 30, 0, 0)
    dG, dgs, do = dst_map if dst_map else (1 << 30, 0, 0)
    _C.check(_C.lib().gpvb200_copy_rows(_C.ptr(_req(src, BF16)), ctypes.c_int64(src.stride(0)), sG, sgs, so, _C.ptr(_req(dst, BF16)),
                                        ctypes.c_int64(dst.stride(0)), dG, dgs, do, ctypes.c_int64(M), D, _C.stream_ptr()),
             "copy_rows")


def cast_bf16(src, out=None):
    y = out if out is not None else torch.empty(src.shape, device=src.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_cast_f32_bf16(_C.ptr(_req(src, torch.float32)), _C.ptr(y), ctypes.c_int64(src.numel()),
                                            _C.stream_ptr()), "cast")
    return y


def add(a, b, out=None):
    """out = a + b (bf16 2-D, row strides allowed)."""
    M, D = a.shape
    y = out if out is not None else torch.empty((M, D), device=a.device, dtype=BF16)
    _C.check(_C.lib().gpvb200_add_bf16(_C.ptr(_req(a, BF16)), ctypes.c_int64(a.stride(0)), _C.ptr(_req(b, BF16)),
                                       ctypes.c_int64(b.stride(0)), _C.ptr(y), ctypes.c_int64(y.stride(0)), ctypes.c_int64(M), D,
                                       _C.stream_ptr()), "add_bf16")
    return y


def unpack_conv_grad(src, dst, accumulate=False):
    """src fp32 [taps,O,I] -> dst fp32 [O,I,kh,kw] (Conv2d weight layout)."""
    taps, O, I = src.shape
    _C.check(_C.lib().gpvb200_unpack_conv_grad(_C.ptr(_req(src, torch.float32)), _C.ptr(_req(dst, torch.float32)), O, I, taps,
                                               int(accumulate), _C.stream_ptr()), "unpack_conv_grad")


class PackPlan:
    """One-launch multi-tensor fp32 -> bf16 weight packing (gpvb200_pack_weights).  items: list of
    (src fp32 tensor, dst bf16 tensor, scale fp32 tensor or None, O, I, taps, mode)."""

    def __init__(self, items, device):
        import numpy as np
        chunk = _C.lib().gpvb200_pack_chunk()
        self.keep = items
        rec = np.zeros(len(items), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("scale", "<u8"), ("O", "<i4"), ("I", "<i4"),
                                                   ("taps", "<i4"), ("mode", "<i4")]))
        assert rec.dtype.itemsize == _C.lib().gpvb200_pack_item_size()
        bi, bc = [], []
        for i, (src, dst, scale, O, I, taps, mode) in enumerate(items):
            _req(src, torch.float32), _req(dst, BF16)
            rec[i] = (src.data_ptr(), dst.data_ptr(), scale.data_ptr() if scale is not None else 0, O, I, taps, mode)
            n = O * 152 if mode == 1 else taps * O * I
            assert dst.numel() >= n and src.is_contiguous() and dst.is_contiguous()
            nb = (n + chunk - 1) // chunk
            bi += [i] * nb
            bc += list(range(nb))
        self.items = torch.from_numpy(rec.view(np.uint8).copy()).to(device)
        self.blk_item = torch.tensor(bi, dtype=torch.int32).to(device)
        self.blk_chunk = torch.tensor(bc, dtype=torch.int32).to(device)

    def run(self):
        _C.check(_C.lib().gpvb200_pack_weights(_C.ptr(self.items), _C.ptr(self.blk_item), _C.ptr(self.blk_chunk),
                                               self.blk_item.numel(), _C.stream_ptr()), "pack_weights")


def bn_fold(w, b, rm, rv, scale, bias):
    _C.check(_C.lib().gpvb200_bn_fold(_C.ptr(w), _C.ptr(b), _C.ptr(rm), _C.ptr(rv), _C.ptr(scale), _C.ptr(bias), w.numel(),
                                      _C.stream_ptr()), "bn_fold")


# ------------------------------------------------------------------------------------------------ criterion
def ce_fwd_bwd(logits, targets, row_weight, loss_sum, dlogits=None, row_loss=None):
    rows, V = logits.shape
    _C.check(_C.lib().gpvb200_ce_fwd_bwd(_C.ptr(_req(logits, torch.float32)), ctypes.c_int64(logits.stride(0)),
                                         _C.ptr(_req(targets, torch.int64)), _C.ptr(_req(row_weight, torch.float32)),
                                         _C.ptr(_req(loss_sum, torch.float32)), _C.ptr(row_loss), _C.ptr(dlogits),
                                         ctypes.c_int64(dlogits.stride(0) if dlogits is not None else 0), rows, V,
                                         _C.stream_ptr()), "ce_fwd_bwd")


def set_criterion(logits, boxes, tgt_boxes, tgt_offsets, idx_q, idx_t, loc_valid, *, eos_coef, weight_sum, num_boxes,
                  wt_ce, wt_bbox, wt_giou, out3, dlogits, dbox_pre):
    B, Q = loc_valid.shape[0], logits.shape[0] // loc_valid.shape[0]
    Kmax = idx_q.shape[1] if idx_q is not None else 0
    f = ctypes.c_float
    _C.check(_C.lib().gpvb200_set_criterion(
        _C.ptr(_req(logits, torch.float32)), ctypes.c_int64(logits.stride(0)), _C.ptr(_req(boxes, torch.float32)),
        ctypes.c_int64(boxes.stride(0)), _C.ptr(tgt_boxes), _C.ptr(_req(tgt_offsets, torch.int32)), _C.ptr(idx_q), _C.ptr(idx_t),
        Kmax, _C.ptr(_req(loc_valid, torch.uint8)), B, Q, f(eos_coef), f(weight_sum), f(num_boxes), f(wt_ce), f(wt_bbox),
        f(wt_giou), _C.ptr(_req(out3, torch.float32)), _C.ptr(_req(dlogits, torch.float32)), _C.ptr(_req(dbox_pre, BF16)),
        ctypes.c_int64(dbox_pre.stride(0)), _C.stream_ptr()), "set_criterion")
