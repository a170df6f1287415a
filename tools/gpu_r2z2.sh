#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r2z_bench_$label.json 2>> $out/r2z_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2z_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$label failed", e)
PY
}
run fused1 X=1
run unfused1 GPVB200_MLP_BWD=0
run fused2 X=1
run unfused2 GPVB200_MLP_BWD=0
run fused_nolazy GPVB200_LAZY_JOIN=0
