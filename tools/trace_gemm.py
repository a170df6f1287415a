"""Developer tool: clock64 timeline of CTA 0 of the contraction kernel (producer warps / MMA issuer / first and last epilogue warp)
on the narrow (N = 64) shapes of the frozen trunk prefix, through a SECOND library built with -DGPV_GEMM_TRACE (the shipped
libgpvb200.so carries no trace code).

    python tools/trace_gemm.py --build          # anywhere: nvcc cross-compiles lib/libgpvb200_trace.so
    python tools/trace_gemm.py [shape ...]       # GPU box: prints the per-tile timeline (clocks relative to the first stamp)
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpv-1_b200")
TRACE_SO = os.path.join(PKG, "lib", "libgpvb200_trace.so")
N_TR = 96


def build():
    sys.path.insert(0, PKG)
    import build as b   # noqa: E402  (flags of the shipped build)
    objdir = os.path.join(PKG, "lib", "obj_trace")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu"))):
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([b.NVCC] + b.FLAGS + ["-DGPV_GEMM_TRACE", "-c", s, "-o", o]))
    if any(p.wait() for p in procs):
        raise SystemExit("trace build failed")
    subprocess.check_call([b.NVCC, "-arch=sm_100a", "-shared", "-o", TRACE_SO] + objs + ["-lcudart"])
    print(TRACE_SO)


def main(which):
    sys.path.insert(0, ROOT)
    import torch
    from gpv1_b200 import _C
    time_only = "--time-only" in sys.argv     # the shipped library, no stamps: the numbers that count
    if not time_only:
        _C.SO_PATH = TRACE_SO
    from gpv1_b200 import kernels as k
    dev = torch.device("cuda:0")
    BF = torch.bfloat16
    L = _C.lib()

    def lin(M, N, K, res=False):
        x = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
        b = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).to(BF) if res else None
        y = torch.empty(M, N, device=dev, dtype=BF)
        return lambda: k.linear(x, w, b, act=k.ACT_RELU, residual=r, out=y)

    def conv3(C=64, B=32, H=120, W=160):
        x = torch.randn(B, H, W, C, device=dev).to(BF)
        w = (torch.randn(9, C, C, device=dev) / (9 * C) ** 0.5).to(BF)
        b = torch.randn(C, device=dev)
        return lambda: k.conv(x, w, ksize=3, bias=b, act=k.ACT_RELU)

    def stem(B=32, H=480, W=640):
        img = torch.randint(0, 255, (B, H, W, 3), device=dev, dtype=torch.uint8)
        xv, Ho, Wo = k.stem_s2d(img)
        w = (torch.randn(4, 64, 64, device=dev) / 14).to(BF)
        b = torch.randn(64, device=dev)
        return lambda: k.conv(xv, w, ksize=7, taps=k.STEM_TAPS, Ho=Ho, Wo=Wo, N=64, K=64, bias=b, act=k.ACT_RELU)

    def conv3n(C, H, W, B=32):
        x = torch.randn(B, H, W, C, device=dev).to(BF)
        w = (torch.randn(9, C, C, device=dev) / (9 * C) ** 0.5).to(BF)
        b = torch.randn(C, device=dev)
        return lambda: k.conv(x, w, ksize=3, bias=b, act=k.ACT_RELU)

    def dgrad3(C, H, W, B=32):
        from gpv1_b200.convops import conv_dgrad
        dy = torch.randn(B, H, W, C, device=dev).to(BF)
        w = (torch.randn(9, C, C, device=dev) / (9 * C) ** 0.5).to(BF)
        aux = torch.randn(B, H, W, C, device=dev).to(BF)
        return lambda: conv_dgrad(dy, w, ksize=3, stride=1, in_hw=(H, W), aux=aux, aux_mode=k.AUX_RELU_MASK)

    def wgrad3(C, H, W, B=32):
        dy = torch.randn(B, H, W, C, device=dev).to(BF)
        x = torch.randn(B, H, W, C, device=dev).to(BF)
        dw = torch.zeros(9, C, C, device=dev)
        return lambda: k.conv_wgrad(dy, x, dw, ksize=3)

    def wgrad1(M, N, K):
        dy = torch.randn(M, N, device=dev).to(BF)
        x = torch.randn(M, K, device=dev).to(BF)
        dw = torch.zeros(N, K, device=dev)
        return lambda: k.linear_wgrad(dy, x, dw)

    def lin_plain(M, N, K):
        x = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
        b = torch.randn(N, device=dev)
        return lambda: k.linear(x, w, b)

    shapes = {
        "l1.conv1.k256": lambda: lin(614400, 64, 256),
        "l1.conv1.k64": lambda: lin(614400, 64, 64),
        "l1.conv3": lambda: lin(614400, 256, 64, res=True),
        "l1.conv2": conv3,
        "stem": stem,
        "l3.conv1": lambda: lin(38400, 256, 1024),
        "l2.conv2": lambda: conv3n(128, 60, 80),
        "l3.conv2": lambda: conv3n(256, 30, 40),
        "l3.conv2.dgrad": lambda: dgrad3(256, 30, 40),
        "l3.conv2.wgrad": lambda: wgrad3(256, 30, 40),
        "l3.conv3.wgrad": lambda: wgrad1(38400, 1024, 256),
        "l2.conv1.wgrad": lambda: wgrad1(153600, 128, 512),
        "enc.qk": lambda: lin_plain(9600, 512, 256),
        "txt.qkv": lambda: lin_plain(640, 2304, 768),
    }
    for name in (which or list(shapes)):
        fn = shapes[name]()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 5
        if time_only:
            print(f"== {name}: {us:.1f} us per launch (shipped library)")
            continue
        tr = torch.zeros(6 * N_TR * 8, dtype=torch.int64, device=dev)
        _C.check(L.gpvb200_gemm_trace(ctypes.c_void_p(tr.data_ptr())), "gemm_trace")
        fn()
        torch.cuda.synchronize()
        _C.check(L.gpvb200_gemm_trace(ctypes.c_void_p(0)), "gemm_trace")
        t = tr.cpu().view(6, N_TR, 8)
        t0 = int(t[t > 0].min())
        rel = lambda v: int(v) - t0 if int(v) > 0 else -1   # noqa: E731
        print(f"== {name}: {us:.1f} us per launch.  clocks of CTA 0 relative to its first stamp")
        print("   P[stage use] = [slot free, loads issued];  M[item] = [top, acc free, first stage landed, last stage landed, committed];"
              "  E / E'[item] = first / last epilogue warp [top, before acc wait, acc full, tmem read, stores issued, released]")
        print("   stage use: producer [slot free, loads issued]   MMA warp [stage landed (seen), MMAs + commit issued]")
        for g in range(16, 64):
            if int(t[0, g, 0]) > 0:
                print(f"   use {g}: P", [rel(v) for v in t[0, g, :2]], " M", [rel(v) for v in t[1, g, :2]])
        for j in range(24):
            if int(t[3, j, 0]) <= 0:
                break
            print(f"   item {j}: M", [rel(v) for v in t[3, j, :5]], " E", [rel(v) for v in t[4, j, :6]], " E'", [rel(v) for v in t[5, j, :6]])
        # steady-state period per work item
        tops = [int(t[3, j, 4]) for j in range(N_TR) if int(t[3, j, 4]) > 0]
        if len(tops) > 12:
            print(f"   steady-state period: {(tops[-1] - tops[8]) / (len(tops) - 9):.0f} clocks per work item ({len(tops)} items traced)")


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main([a for a in sys.argv[1:] if not a.startswith("-")])
