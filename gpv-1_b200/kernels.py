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


def _req(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("gpvb200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return t


def _launch_gemm(d: _C.GemmDesc):
    _C.check(_C.lib().gpvb200_gemm(ctypes.byref(d), _C.stream_ptr()), "gemm")


def gemm(A, B, D, *, M, N, K, lda, ldb, ldd, a_mn=False, b_mn=False, batch=1, a_bs=0, b_bs=0, d_bs=0,
         bias=None, rowscale=None, residual=None, ldr=0, aux=None, ldaux=0, aux_mode=AUX_NONE, act=ACT_NONE,
         alpha=1.0, atomic=False, splits=1, D2=None):
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
    d.residual, d.aux = _C.ptr(_req(residual, BF16)), _C.ptr(_req(aux, BF16))
    d.lda, d.ldb, d.ldd, d.ldr, d.ldaux = lda, ldb, ldd, ldr, ldaux
    d.a_batch_stride, d.b_batch_stride, d.d_batch_stride = a_bs, b_bs, d_bs
    _launch_gemm(d)
    return D


def linear(x, w, bias=None, *, act=ACT_NONE, residual=None, out_dtype=BF16, out=None, out2=None, alpha=1.0):
    """y[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias + residual).  x, w bf16 row-major (last dim contiguous)."""
    _req(x, BF16), _req(w, BF16)
    M, K = x.shape
    N = w.shape[0]
    y = out if out is not None else torch.empty((M, N), device=x.device, dtype=out_dtype)
    return gemm(x, w, y, M=M, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldd=y.stride(0), bias=bias,
                residual=residual, ldr=residual.stride(0) if residual is not None else 0, act=act, D2=out2, alpha=alpha)


def linear_dgrad(dy, w, *, aux=None, aux_mode=AUX_NONE, residual=None, out=None, out_dtype=BF16):
    """dx[M,K] = (dy[M,N] @ w[N,K] + residual) (* mask(aux)).  w is read in its forward layout (MN-major B)."""
    _req(dy, BF16), _req(w, BF16)
    M, N = dy.shape
    K = w.shape[1]
    dx = out if out is not None else torch.empty((M, K), device=dy.device, dtype=out_dtype)
    return gemm(dy, w, dx, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldd=dx.stride(0), b_mn=True,
                residual=residual, ldr=residual.stride(0) if residual is not None else 0,
                aux=aux, ldaux=aux.stride(0) if aux is not None else 0, aux_mode=aux_mode)


def linear_wgrad(dy, x, dw, *, rowscale=None, splits=0):
    """dw[N,K] (fp32) += dy[M,N]^T @ x[M,K]; both operands read in place (MN-major), split over M."""
    _req(dy, BF16), _req(x, BF16), _req(dw, torch.float32)
    M, N = dy.shape
    K = x.shape[1]
    if splits <= 0:
        tiles = ((N + 127) // 128) * ((K + 127) // 128)
        splits = max(1, min((M + 63) // 64, (2 * 148 + tiles - 1) // tiles))
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
def matcher_cost(logits, boxes, tgt_boxes, tgt_labels, tgt_offsets, Tmax, w_class, w_bbox, w_giou):
    """fp32 cost blocks [B,Q,Tmax] (utils/matcher.py:53-72)."""
    B, Q, C = logits.shape
    cost = torch.zeros((B, Q, max(Tmax, 1)), device=logits.device, dtype=torch.float32)
    if B and Tmax:
        _C.check(_C.lib().gpvb200_matcher_cost(
            _C.ptr(_req(logits, torch.float32)), _C.ptr(_req(boxes, torch.float32)), _C.ptr(_req(tgt_boxes, torch.float32)),
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
