#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py tests/test_blocks_gpu.py -x -q -m gpu > $out/r3a_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3a_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r3a_bench_$label.json 2>> $out/r3a_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r3a_bench_$label.json")); a=d["attention_kernel"]["bwd_dh32"]; print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "bwd_dh32 avg us", round(a["avg_launch_us"],1))
except Exception as e: print("$label failed", e)
PY
}
run nt320 X=1
run nt256 GPVB200_ATTN_BWD_320=0
run nt320b X=1
