"""Developer tool (GPU): the production contraction kernel with and without the CTA-pair variant on the step's deep shapes --
plain GEMMs, 3x3 / 1x1 implicit-GEMM convolutions (forward and data gradient) and weight gradients, with the epilogues the
engine uses -- side by side.  GPVB200_PAIR is read once per process, so the script runs itself once per setting and joins the
tables.  Each timing replays a CUDA graph of `reps` back-to-back launches (no host gaps), after a warm-up.

    python tools/prof_pair.py                 # B200; prints   shape | single us | pair us | ratio | pair launches
    python tools/prof_pair.py --child         # one setting (used internally): JSON lines on stdout
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (tag, kind, args)   kinds: lin = y = act(x W^T + b [+ r]); dgrad = dx = dy W; wgrad = dW += dy^T x; conv / conv_dgrad: NHWC
CASES = [
    ("l3.conv1 1x1 38400x256x1024 relu", "lin", (38400, 256, 1024, False)),
    ("l3.conv3.dgrad 38400: 1024->256", "dgrad", (38400, 1024, 256)),
    ("l4.conv1 1x1 9600x512x2048 relu", "lin", (9600, 512, 2048, False)),
    ("l4.conv3.dgrad 9600: 2048->512", "dgrad", (9600, 2048, 512)),
    ("enc.ffn2 9600x256x2048 +res", "lin", (9600, 256, 2048, True)),
    ("enc.ffn1.dgrad 9600: 2048->256", "dgrad", (9600, 2048, 256)),
    ("l2.conv2 3x3 32x60x80 128->128", "conv", (32, 60, 80, 128, 128, 3, 1)),
    ("l3.conv2 3x3 32x30x40 256->256", "conv", (32, 30, 40, 256, 256, 3, 1)),
    ("l4.conv2 3x3 32x15x20 512->512", "conv", (32, 15, 20, 512, 512, 3, 1)),
    ("l3.conv2.dgrad 3x3 32x30x40 256->256", "conv_dgrad", (32, 30, 40, 256, 256, 3, 1)),
    ("l4.conv2.dgrad 3x3 32x15x20 512->512", "conv_dgrad", (32, 15, 20, 512, 512, 3, 1)),
    ("l3.conv3.wgrad 38400: 1024x256", "wgrad", (38400, 1024, 256)),
    ("l3.conv1.wgrad 38400: 256x1024", "wgrad", (38400, 256, 1024)),
    ("l4.conv3.wgrad 9600: 2048x512", "wgrad", (9600, 2048, 512)),
    ("enc.ffn1.wgrad 9600: 2048x256", "wgrad", (9600, 2048, 256)),
]


def child(reps=10):
    import torch
    from gpv1_b200 import _C, convops
    from gpv1_b200 import kernels as k
    dev = torch.device("cuda:0")
    lib = _C.lib()
    BF = torch.bfloat16
    torch.manual_seed(0)

    def bf(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev) * scale).to(BF)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n0 = lib.gpvb200_gemm_pair_launches()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        n_pair = lib.gpvb200_gemm_pair_launches() - n0
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / (3 * reps), n_pair // reps

    for tag, kind, a in CASES:
        if kind == "lin":
            M, N, K, res = a
            x, w, b = bf(M, K), bf(N, K, scale=K ** -0.5), torch.randn(N, device=dev)
            r = bf(M, N) if res else None
            y = torch.empty(M, N, device=dev, dtype=BF)
            fn = lambda: k.linear(x, w, b, act=k.ACT_NONE if res else k.ACT_RELU, residual=r, out=y)   # noqa: E731
            flop = 2.0 * M * N * K
        elif kind == "dgrad":
            M, N, K = a
            dy, w, h = bf(M, N), bf(N, K, scale=N ** -0.5), bf(M, K)
            dx = torch.empty(M, K, device=dev, dtype=BF)
            fn = lambda: k.linear_dgrad(dy, w, aux=h, aux_mode=k.AUX_RELU_MASK, out=dx)   # noqa: E731
            flop = 2.0 * M * N * K
        elif kind == "wgrad":
            M, N, K = a
            dy, x = bf(M, N), bf(M, K)
            dw = torch.zeros(N, K, device=dev)
            fn = lambda: k.linear_wgrad(dy, x, dw)   # noqa: E731
            flop = 2.0 * M * N * K
        else:
            n, H, W, Cin, Cout, ks, stride = a
            x, w, b = bf(n, H, W, Cin), bf(ks * ks, Cout, Cin, scale=(Cin * ks * ks) ** -0.5), torch.randn(Cout, device=dev)
            flop = 2.0 * n * H * W * Cin * Cout * ks * ks
            if kind == "conv":
                y = torch.empty(n, H, W, Cout, device=dev, dtype=BF)
                fn = lambda: k.conv(x, w, ksize=ks, stride=stride, bias=b, act=k.ACT_RELU, out=y)   # noqa: E731
            else:
                dy, h = bf(n, H, W, Cout), bf(n, H, W, Cin)
                dx = torch.empty(n, H, W, Cin, device=dev, dtype=BF)
                fn = lambda: convops.conv_dgrad(dy, w, ksize=ks, stride=stride, in_hw=(H, W), aux=h, aux_mode=k.AUX_RELU_MASK, out=dx)   # noqa: E731
        us, n_pair = timed(fn)
        print(json.dumps({"tag": tag, "us": us, "tflops": flop / us / 1e6, "pair_launches": n_pair}), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    settings = [("single", {"GPVB200_PAIR": "0"}), ("pair", {"GPVB200_PAIR": "8"})]
    rows = {}
    for name, env in settings:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, **env), capture_output=True, text=True)
        if res.returncode != 0:
            print(f"[{name}] failed:\n{res.stderr[-2000:]}")
        for line in res.stdout.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                rows.setdefault(d["tag"], {})[name] = d
    print(f"{'shape':44s} {'single us':>10s} {'TFLOP/s':>8s} {'pair us':>9s} {'ratio':>6s} {'pair+wgrad us':>14s} {'ratio':>6s}  pair launches")
    for tag, _, _ in CASES:
        r = rows.get(tag, {})
        s, p, q = r.get("single"), r.get("pair"), r.get("pair+wgrad")
        if not s:
            continue
        cell = lambda d: (f"{d['us']:9.1f} {s['us'] / d['us']:6.2f}" if d else f"{'-':>9s} {'-':>6s}")   # noqa: E731
        print(f"{tag:44s} {s['us']:10.1f} {s['tflops']:8.1f} {cell(p)} {cell(q):>21s}  {p['pair_launches'] if p else '-'} / {q['pair_launches'] if q else '-'}")


if __name__ == "__main__":
    main()
