"""Developer tool (GPU): per-kernel TIMELINE of the captured training step under CUDA-graph replay, through torch.profiler
(CUPTI activity records: start, duration, stream, grid of every kernel; no replay, no serialisation, unlike ncu), and
what it says about the things a launch list cannot show: how much of the step the SMs sit idle, how much of it runs
more than one kernel at a time (lanes), how long the tails of the persistent GEMM launches are.

    python tools/profile_step.py [--batch 32] [--steps 3] [--out gpurun_out/rX_timeline]      # B200
    python tools/profile_step.py --analyse trace.json                                        # anywhere: re-read a saved trace

Writes <out>.json (summary) and <out>_trace.json.gz (chrome trace).  Numbers taken under the profiler are for analysis only,
never bench values.
"""
import argparse
import collections
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_SM = 148


def kernel_events(trace):
    """Kernel records of a chrome trace written by torch.profiler: (name, start us, duration us, stream, CTAs)."""
    out = []
    for e in trace.get("traceEvents", []):
        if e.get("ph") != "X" or e.get("cat") not in ("kernel", "Kernel"):
            continue
        a = e.get("args", {})
        grid = a.get("grid", [1, 1, 1])
        ctas = int(grid[0]) * int(grid[1]) * int(grid[2]) if isinstance(grid, (list, tuple)) and len(grid) == 3 else 1
        out.append((e["name"], float(e["ts"]), float(e["dur"]), a.get("stream", 0), ctas))
    out.sort(key=lambda r: r[1])
    return out


def split_steps(ev, marker="stem_s2d"):
    """One step = the records between two consecutive stem kernels (BERT's first kernels, enqueued on their own lane ahead of
    the stem, are attributed to the previous step: a constant offset that does not change any total)."""
    idx = [i for i, r in enumerate(ev) if marker in r[0]]
    return [(ev[a:b], ev[b][1]) for a, b in zip(idx[:-1], idx[1:])]


def analyse_step(step, t_next):
    """Timeline statistics of one step (it ends where the next step's stem kernel starts).  SM occupancy is estimated from the
    grid: a kernel with g CTAs holds min(g, 148) SMs for its whole duration (the persistent GEMM and the attention kernels run
    one or two CTAs per SM), capped at 148 in total."""
    t0 = min(r[1] for r in step)
    t1 = max(t_next, max(r[1] + r[2] for r in step))
    points = []
    for name, ts, dur, stream, ctas in step:
        sms = min(ctas, N_SM)
        points.append((ts, 1, sms))
        points.append((ts + dur, -1, -sms))
    points.append((t1, 0, 0))
    points.sort()
    busy_sm_time = idle_time = multi_time = 0.0
    live = sms_live = 0
    prev = t0
    for t, dk, dsm in points:
        span = t - prev
        if span > 0:
            busy_sm_time += span * min(sms_live, N_SM)
            if live == 0:
                idle_time += span
            if live >= 2:
                multi_time += span
        live += dk
        sms_live += dsm
        prev = t
    per = collections.OrderedDict()
    for name, ts, dur, stream, ctas in step:
        key = name.split("(")[0][:80]
        c = per.setdefault(key, [0, 0.0, 0])
        c[0] += 1
        c[1] += dur
        c[2] += ctas
    wall = t1 - t0
    small = [r for r in step if r[4] < N_SM // 2]
    return {"wall_us": wall, "kernels": len(step), "kernel_time_us": sum(r[2] for r in step),
            "no_kernel_running_us": idle_time, "two_or_more_kernels_us": multi_time,
            "sm_occupancy_estimate": busy_sm_time / (wall * N_SM) if wall > 0 else 0.0,
            "launches_under_half_the_sms": len(small), "their_time_us": sum(r[2] for r in small),
            "streams": len({r[3] for r in step}),
            "by_kernel": {k: {"launches": v[0], "us": round(v[1], 1), "avg_ctas": round(v[2] / v[0], 1)}
                          for k, v in sorted(per.items(), key=lambda x: -x[1][1])[:25]}}


def analyse(trace):
    ev = kernel_events(trace)
    steps = split_steps(ev)
    if not steps:
        return {"error": "no complete step between two stem kernels in the trace", "kernel_records": len(ev)}
    res = [analyse_step(s, t_next) for s, t_next in steps]
    res.sort(key=lambda r: r["wall_us"])
    return {"steps_in_trace": len(steps), "median_step": res[len(res) // 2]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline"))
    ap.add_argument("--analyse", default=None, help="re-analyse a saved chrome trace (.json or .json.gz) instead of running")
    args = ap.parse_args()
    if args.analyse:
        op = gzip.open if args.analyse.endswith(".gz") else open
        with op(args.analyse, "rt") as f:
            print(json.dumps(analyse(json.load(f)), indent=1))
        return
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    from gpv1_b200.config import load_config
    from gpv1_b200.model import GPV
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    model = GPV(load_config().model, vocab=bench.vocab_list(bench.V_BENCH), seed=0).to(dev)
    model.train()
    images, qids, ans, targets = bench.make_batch(args.batch, seed=1000)
    b = (images.to(dev), qids.to(dev), ans.to(dev), [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in t.items()} for t in targets])
    model.capture_step(*b)

    def step():
        model(*b).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps + 1):              # n + 1 stem kernels bound n complete steps
            step()
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    raw = args.out + "_trace.json"
    prof.export_chrome_trace(raw)
    with open(raw) as f:
        trace = json.load(f)
    with gzip.open(raw + ".gz", "wt") as f:
        json.dump(trace, f)
    os.remove(raw)
    res = analyse(trace)
    with open(args.out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:4000])


if __name__ == "__main__":
    main()
