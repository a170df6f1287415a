#!/bin/bash
# round-2 GPU call E: multi-producer GEMM -- parity suite, per-shape timings and bench with 1 / 2 / 3 producer warps
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2e_pytest.log
for np in 1 2 3; do
  GPVB200_PRODUCERS=$np timeout 300 python tools/prof_gemm.py > $out/r2e_gemm_shapes_p$np.txt 2>&1
  GPVB200_PRODUCERS=$np timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2e_bench_p$np.json 2>> $out/r2e_bench.err
  python - <<PY
import json
d=json.load(open("$out/r2e_bench_p$np.json")); print("producers=$np", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s", "roofline", round(d["roofline"]["frac"],3))
PY
done
paste -d'|' $out/r2e_gemm_shapes_p1.txt $out/r2e_gemm_shapes_p3.txt | cut -c1-200
