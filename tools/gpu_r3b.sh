#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py tests/test_blocks_gpu.py tests/test_model_gpu.py -x -q -m gpu > $out/r3b_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3b_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r3b_bench_$label.json 2>> $out/r3b_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r3b_bench_$label.json")); a=d["attention_kernel"]; print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), {k: round(v["avg_launch_us"],1) for k,v in a.items() if isinstance(v,dict) and k.startswith("bwd")})
except Exception as e: print("$label failed", e)
PY
}
run a X=1
run b X=1
