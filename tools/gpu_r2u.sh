#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2u_bench_$label.json 2>> $out/r2u_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2u_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$label failed", e)
PY
}
run wg32 GPVB200_WGRAD_CTAS=32
run wg40 GPVB200_WGRAD_CTAS=40
run wg48 GPVB200_WGRAD_CTAS=48
run wg56 GPVB200_WGRAD_CTAS=56
run wg64 GPVB200_WGRAD_CTAS=64
run wg64_lanes2 GPVB200_WGRAD_CTAS=64 GPVB200_WGRAD_LANES=2
run wg48_lanes4 GPVB200_WGRAD_CTAS=48 GPVB200_WGRAD_LANES=4
run wg64_kper16 GPVB200_WGRAD_CTAS=64 GPVB200_MIN_KPER=16
