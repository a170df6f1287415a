#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py tests/test_blocks_gpu.py -x -q -m gpu > $out/r2v_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2v_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2v_bench_$label.json 2>> $out/r2v_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2v_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$label failed", e)
PY
}
run base X=1
run base2 X=2
