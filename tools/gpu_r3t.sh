#!/bin/bash
# lane CTA cap / lane count again, now that the chain waits for the lanes only at the end of backward (N = 1)
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r3t_bench_$label.json 2>> $out/r3t_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r3t_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$label failed", e)
PY
}
run cap74 X=1
run cap56 GPVB200_WGRAD_CTAS=56
run cap96 GPVB200_WGRAD_CTAS=96
run cap120 GPVB200_WGRAD_CTAS=120
run lanes2 GPVB200_WGRAD_LANES=2
run lanes4 GPVB200_WGRAD_LANES=4
run cap96_lanes2 GPVB200_WGRAD_CTAS=96 GPVB200_WGRAD_LANES=2
