#!/bin/bash
# round-2 GPU call H: tile-width sweep with the new issue paths, pair variant inside the step, per-entry breakdown
out=gpurun_out; mkdir -p $out
timeout 900 python tools/sweep_bn.py profiles/r1F_breakdown_by_shape.json > $out/r2h_sweep_bn.txt 2>&1; echo "sweep exit $?"; grep -c "model misses" $out/r2h_sweep_bn.txt; tail -1 $out/r2h_sweep_bn.txt
for thr in 8 16; do
  GPVB200_PAIR=$thr timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2h_bench_pair$thr.json 2>> $out/r2h_bench.err
  python - <<PY
import json
d=json.load(open("$out/r2h_bench_pair$thr.json")); print("pair>=$thr", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s")
PY
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown $out/r2h_breakdown.json > $out/r2h_bench_eager.json 2>> $out/r2h_bench.err
echo "breakdown exit $?"
