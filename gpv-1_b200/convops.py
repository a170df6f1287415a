"""Convolution data-gradient built from the implicit-GEMM kernel (mode 1 of gpvb200_gemm).

Forward (backbone.py:72, torchvision Bottleneck): y[ho,wo] = sum_t x[ho*s+dh_t, wo*s+dw_t] W_t.
Data gradient: dx[hi,wi] = sum_t dy[(hi-dh_t)/s, (wi-dw_t)/s] W_t^T over the taps where the division is exact.
For s=1 that is one launch with mirrored taps; for s=2 one launch per output parity class (hi%2, wi%2), each
reading dy at unit stride and scattering into its interleaved positions of dx.
"""
import torch

from . import kernels as k


def conv_dgrad(dy, w, *, ksize, stride, in_hw, aux=None, aux_mode=k.AUX_NONE, residual=None, out=None, only_parity=None):
    """dy [n,Ho,Wo,Cout] bf16, w [taps,Cout,Cin] bf16 (forward layout) -> dx [n,Hi,Wi,Cin] bf16.

    Epilogue: dx = (acc + residual) * mask(aux), with residual/aux indexed like dx.  only_parity = (ph, pw) (stride 2): only that
    parity class of dx is written (the caller owns the other positions; residual may alias out for an in-place accumulation).
    """
    n, Ho, Wo, Cout = dy.shape
    Hi, Wi = in_hw
    Cin = w.shape[2]
    taps = [(0, 0)] if ksize == 1 else [(r - 1, s - 1) for r in range(3) for s in range(3)]
    if stride == 1:
        return k.conv(dy, w, ksize=ksize, stride=1, taps=[(-dh, -dw) for dh, dw in taps], b_mn=True, N=Cin, K=Cout,
                      aux=aux, aux_mode=aux_mode, residual=residual, out=out, Ho=Hi, Wo=Wi)
    assert stride == 2
    dx = out if out is not None else torch.empty((n, Hi, Wi, Cin), device=dy.device, dtype=torch.bfloat16)
    for ph in range(2):
        for pw in range(2):
            if only_parity is not None and (ph, pw) != tuple(only_parity):
                continue
            sel = [(i, (ph - dh) // 2, (pw - dw) // 2) for i, (dh, dw) in enumerate(taps)
                   if (ph - dh) % 2 == 0 and (pw - dw) % 2 == 0]
            nh, nw = (Hi - ph + 1) // 2, (Wi - pw + 1) // 2
            if nh <= 0 or nw <= 0:
                continue
            if not sel:
                # no tap lands on this parity class (1x1 stride-2): gradient is just the residual path
                sub = dx[:, ph::2, pw::2, :]
                if residual is not None:
                    r = residual[:, ph::2, pw::2, :].float()
                    if aux_mode == k.AUX_RELU_MASK:
                        r = r * (aux[:, ph::2, pw::2, :] > 0)
                    sub.copy_(r)
                else:
                    sub.zero_()
                continue
            k.conv(dy, w, ksize=ksize, stride=1, taps=[(a, b) for _, a, b in sel], tap_w=[i for i, _, _ in sel],
                   b_mn=True, N=Cin, K=Cout, aux=aux, aux_mode=aux_mode, residual=residual, out=dx, Ho=nh, Wo=nw,
                   out_geom=(Hi, Wi, 2, ph, pw))
    return dx
