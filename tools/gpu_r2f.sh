#!/bin/bash
# round-2 GPU call F: elect.sync issue paths + multi-producer GEMM -- parity suite, per-shape timings, bench A/B, mlp trace
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2f_pytest.log
timeout 120 python tools/trace_layer.py 9600 > $out/r2f_trace.txt 2>&1; sed -n 12,16p $out/r2f_trace.txt; tail -2 $out/r2f_trace.txt
timeout 200 python tools/prof_layer.py > $out/r2f_prof_layer.txt 2>&1; cat $out/r2f_prof_layer.txt
for np in 1 3; do
  GPVB200_PRODUCERS=$np timeout 300 python tools/prof_gemm.py > $out/r2f_gemm_shapes_p$np.txt 2>&1
  GPVB200_PRODUCERS=$np timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2f_bench_p$np.json 2>> $out/r2f_bench.err
  python - <<PY
import json
d=json.load(open("$out/r2f_bench_p$np.json")); print("producers=$np", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s", "roofline", round(d["roofline"]["frac"],3))
PY
done
paste -d'|' $out/r2f_gemm_shapes_p1.txt $out/r2f_gemm_shapes_p3.txt | cut -c1-200
