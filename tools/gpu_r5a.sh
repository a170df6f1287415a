#!/bin/bash
# re-scan of the tile-width cost model / split-K floor / lane cap after the round-4 main-loop changes (resident step only, N = 1)
out=gpurun_out; mkdir -p $out; : > $out/r5a_scan.txt
run() { label=$1; shift
  env "$@" timeout 200 python bench.py --profiling --steps 30 --warmup 5 > $out/r5a_$label.json 2>> $out/r5a.err
  python - <<PY | tee -a $out/r5a_scan.txt
import json
try:
    d=json.load(open("$out/r5a_$label.json")); print("$label", round(d["ms_per_step_under_profiler"],3), "ms/step", d["launches_per_step"], "launches")
except Exception as e: print("$label failed", e)
PY
}
run base X=1
run titer_18_18 GPVB200_TITER=0.18,0.0018
run titer_12_20 GPVB200_TITER=0.12,0.0020
run titer_30_18 GPVB200_TITER=0.30,0.0018
run titer_24_12 GPVB200_TITER=0.24,0.0012
run titer_24_25 GPVB200_TITER=0.24,0.0025
run kper4 GPVB200_MIN_KPER=4
run kper12 GPVB200_MIN_KPER=12
run cap56 GPVB200_WGRAD_CTAS=56
run cap96 GPVB200_WGRAD_CTAS=96
run base2 X=1
