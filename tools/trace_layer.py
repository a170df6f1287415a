"""Developer tool (GPU box): clock64 timeline of CTA 0 of mlp_block_fwd (producer / MMA issuer / epilogue warp 2)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpv1_b200 import kernels as k, _C

dev = torch.device("cuda:0")
BF = torch.bfloat16
M, dff = int(sys.argv[1]) if len(sys.argv) > 1 else 9600, 2048
n = dff // 64
x = torch.randn(M, 256, device=dev).to(BF)
w1 = (torch.randn(dff, 256, device=dev) / 16).to(BF)
w2 = (torch.randn(256, dff, device=dev) / 45).to(BF)
b1, b2, g, b = (torch.randn(i, device=dev) for i in (dff, 256, 256, 256))
for _ in range(3):
    k.mlp_block_fwd(x, w1, b1, w2, b2, g, b, 1e-5)
tr = torch.zeros(3 * (n + 1) * 8, dtype=torch.int64, device=dev)
_C.lib().gpvb200_layer_trace(ctypes.c_void_p(tr.data_ptr()))
k.mlp_block_fwd(x, w1, b1, w2, b2, g, b, 1e-5)
torch.cuda.synchronize()
_C.lib().gpvb200_layer_trace(ctypes.c_void_p(0))
t = tr.cpu().view(3, n + 1, 8)
t0 = int(t[t > 0].min())
rel = lambda v: int(v) - t0 if int(v) > 0 else -1
print("producer: per chunk j: [W1(j) slot free, W2(j) slot free]")
print("mma: [W1 landed, acc1 free, F1 issued, W2 landed, H full, F2 issued]")
print("epi(warp 2): [start, acc1 full, ld done, math done, H free, H written, H published]")
for j in range(n + 1):
    print(j, "P", [rel(v) for v in t[0, j, :2]], "M", [rel(v) for v in t[1, j, :6]], "E", [rel(v) for v in t[2, j, :7]])


def attn_trace(B=32, Sq=300, Sk=300):
    qkv = torch.randn(B * Sq, 768, device=dev).to(BF)
    q, kk, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
    for _ in range(3):
        k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=32 ** -0.5)
    tr = torch.zeros(3 * 9 * 8, dtype=torch.int64, device=dev)
    _C.lib().gpvb200_layer_trace(ctypes.c_void_p(tr.data_ptr()))
    k.attn_block_fwd(q, kk, v, B=B, Sq=Sq, Sk=Sk, scale=32 ** -0.5)
    torch.cuda.synchronize()
    _C.lib().gpvb200_layer_trace(ctypes.c_void_p(0))
    t = tr.cpu().view(3, 9, 8)
    t0 = int(t[t > 0].min())
    rel = lambda v: int(v) - t0 if int(v) > 0 else -1
    print("attn_block core, CTA 0.  mma: [loop top, P full, O empty, PV issued]  softmax warp 0: [top, S full, pass 1 done, barrier, O_{h-1} out, pass 2 done, P published]")
    for h in range(8):
        print(h, "M", [rel(v) for v in t[1, h, :4]], "E", [rel(v) for v in t[2, h, :7]])


attn_trace()
