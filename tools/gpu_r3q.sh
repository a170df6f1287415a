#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r3q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3q_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/r3q_bench_$label.json 2>> $out/r3q_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r3q_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "full", round(d["full_step"]["ms_per_step"],3))
except Exception as e: print("$label failed", e)
PY
}
run a X=1
run b X=1
