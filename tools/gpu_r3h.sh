#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_blocks_gpu.py tests/test_model_gpu.py -x -q -m gpu > $out/r3h_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3h_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r3h_bench_$label.json 2>> $out/r3h_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r3h_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "gemm avg us", round(d["roofline"]["avg_launch_us"],2), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print("$label failed", e)
PY
}
run bres X=1
run nobres GPVB200_BRES=0
run bres2 X=1
tail -3 $out/r3h_bench.err
