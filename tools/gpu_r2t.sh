#!/bin/bash
# round-2 GPU call T: lane CTA caps / priorities again, now that the lanes may lag behind the chain
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2t_bench_$label.json 2>> $out/r2t_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2t_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$label failed", e)
PY
}
run base X=1
run wg72 GPVB200_WGRAD_CTAS=72
run wg96 GPVB200_WGRAD_CTAS=96
run wg120 GPVB200_WGRAD_CTAS=120
run prio GPVB200_MAIN_PRIO=1
run prio_wg96 GPVB200_MAIN_PRIO=1 GPVB200_WGRAD_CTAS=96
run lanes4 GPVB200_WGRAD_LANES=4
run lanes2 GPVB200_WGRAD_LANES=2
