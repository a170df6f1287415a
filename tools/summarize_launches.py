"""Developer tool: aggregate an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list over the LAST training step (one stem-to-stem period at the end of the list) into a per-kernel markdown
table; with `--json` also prints the step totals (kernel time, DRAM bytes) as one JSON line; with `--traffic OUT.json`
writes the step totals plus those of the tcgen05 GEMM kernel (what bench.py reads as `roofline.traffic`)."""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, as_json=False, traffic_out=None):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if l.startswith('"')]))
    # one record per launch ID, metrics merged
    launches = collections.OrderedDict()
    for x in rows:
        rec = launches.setdefault(x["ID"], {"name": x["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(x["Metric Value"].replace(",", "")) * UNIT.get(x["Metric Unit"], 1.0)
        if x["Metric Name"].startswith("gpu__time_duration"):
            rec["us"] = v
        elif "bytes_read" in x["Metric Name"]:
            rec["rd"] = v
        elif "bytes_write" in x["Metric Name"]:
            rec["wr"] = v
    recs = list(launches.values())
    idx = [i for i, r in enumerate(recs) if "stem_s2d" in r["name"] or "stem_im2col" in r["name"]]
    # one step = the launches between two consecutive stem kernels (BERT is enqueued ahead of the stem on its own lane)
    step = recs[-(idx[-1] - idx[-2]):] if len(idx) >= 2 else recs
    tot = sum(r["us"] for r in step)
    agg = collections.OrderedDict()
    for r in step:
        n = re.sub(r"\(.*", "", r["name"].replace("(int)", "").replace("(bool)", ""))[:70]
        a = agg.setdefault(n, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["us"]
        a[2] += r["rd"] + r["wr"]
    dram = sum(r["rd"] + r["wr"] for r in step)
    print(f"one step: {len(step)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised), {dram / 1e9:.2f} GB of DRAM traffic\n")
    print("| kernel | launches | us | us/launch | share | DRAM GB |\n|---|---|---|---|---|---|")
    for k, (c, us, by) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if us / tot < 0.001:
            continue
        print(f"| `{k}` | {c} | {us:.0f} | {us / c:.1f} | {100 * us / tot:.1f}% | {by / 1e9:.3f} |")
    if traffic_out:
        g = [r for r in step if "umma_gemm_kernel" in r["name"]]
        gus, gby = sum(r["us"] for r in g), sum(r["rd"] + r["wr"] for r in g)
        import os
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from gpv1_b200 import build
        sha = build.source_sha16()              # csrc/ + include/ + nvcc flags: the code the launch list was taken from
        with open(traffic_out, "w") as f:
            json.dump({"lib_sha16": sha,        # bench.py reports `roofline.traffic` only when this is the library it is running
                       "source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                                 "python bench.py --steps 1 --warmup 1 --profiling --no-graph; last step of the list",
                       "launches_per_step": len(step), "kernel_us_per_step": tot, "dram_bytes_per_step": dram,
                       "gemm_kernel": {"name": "gpv::umma_gemm_kernel<BN,F>", "launches": len(g), "us": gus, "dram_bytes": gby,
                                       "share_of_kernel_time": gus / tot, "dram_bytes_per_launch": gby / max(len(g), 1)}}, f, indent=1)
    if as_json:
        print(json.dumps({"launches": len(step), "kernel_ms": tot / 1e3, "dram_read_bytes": sum(r["rd"] for r in step),
                          "dram_write_bytes": sum(r["wr"] for r in step)}))


if __name__ == "__main__":
    main(sys.argv[1], "--json" in sys.argv, sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None)
