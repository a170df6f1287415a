"""Developer tool (GPU box): time every distinct plain-GEMM shape of the training step (from a bench --breakdown JSON)
with the tile width forced to 64 / 128 / 256 (GPVB200_FORCE_BN, one process per width) and with the cost model's own
choice.  Each shape is timed as 20 back-to-back launches replayed from a CUDA graph (no host launch gaps).

    python tools/sweep_bn.py profiles/r1h_breakdown_by_shape.json            # spawns the 4 runs, prints a table
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def shapes(path):
    d = json.load(open(path))
    out = []
    for key, v in d["entries"].items():
        m = re.match(r"gemm (fwd|dgrad|wgrad) M(\d+) N(\d+) K(\d+) b(\d+) s(\d+)", key)
        if m and int(m.group(5)) == 1:
            out.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), v["calls"]))
    return out


def child(path):
    import torch
    from gpv1_b200 import kernels as k
    dev = torch.device("cuda:0")
    BF = torch.bfloat16
    res = {}
    def mat(r, c, dtype=BF):        # [r, c] view of a buffer whose row stride is a multiple of 8 elements
        return torch.randn(r, (c + 7) // 8 * 8, device=dev).to(dtype)[:, :c]

    for kind, M, N, K, calls in shapes(path):
        torch.manual_seed(0)
        if kind == "fwd":
            x, w, b, y = mat(M, K), mat(N, K), torch.randn((N + 7) // 8 * 8, device=dev)[:N], mat(M, N)
            fn = lambda: k.linear(x, w, b, out=y)
        elif kind == "dgrad":        # dx[M,N] = dy[M,K] @ w[K,N]
            dy, w, dx = mat(M, K), mat(K, N), mat(M, N)
            fn = lambda: k.linear_dgrad(dy, w, out=dx)
        else:                        # dw[M,N] += dy[K,M]^T x[K,N]
            dy, x, dw = mat(K, M), mat(K, N), mat(M, N, torch.float32)
            fn = lambda: k.linear_wgrad(dy, x, dw)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[f"{kind} M{M} N{N} K{K}"] = [1e3 * e0.elapsed_time(e1) / 20, calls]
    print("RESULT " + json.dumps(res))


def main(path):
    runs = {}
    for bn in ("0", "64", "128", "256"):
        env = dict(os.environ, GPVB200_FORCE_BN=bn)
        out = subprocess.run([sys.executable, __file__, "--child", path], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(out.stdout[-2000:], out.stderr[-2000:])
            raise SystemExit(f"child BN={bn} failed")
        runs[bn] = json.loads(line[0][7:])
    tot = {b: 0.0 for b in runs}
    best_tot = 0.0
    print(f"{'shape':34s} calls   model     64    128    256   (us per launch)")
    for key in runs["0"]:
        calls = runs["0"][key][1]
        t = {b: runs[b][key][0] for b in runs}
        for b in runs:
            tot[b] += t[b] * calls
        best_tot += min(t.values()) * calls
        flag = "" if t["0"] <= 1.05 * min(t.values()) else "  <-- model misses"
        print(f"{key:34s} {calls:5d} {t['0']:7.1f} {t['64']:6.1f} {t['128']:6.1f} {t['256']:6.1f}{flag}")
    print("per-step totals (ms): " + ", ".join(f"BN={b}: {v / 1e3:.2f}" for b, v in tot.items()) + f", best-of: {best_tot / 1e3:.2f}")


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main(sys.argv[1])
