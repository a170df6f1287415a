#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_layer_gpu.py tests/test_blocks_gpu.py tests/test_dropout_gpu.py -x -q -rP -m gpu > $out/r2z_pytest.log 2>&1; echo "pytest exit $?"; grep "mlp_block_bwd" $out/r2z_pytest.log | sort -u | head; tail -3 $out/r2z_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2z_bench_$label.json 2>> $out/r2z_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2z_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "loss", d["config"]["loss"])
except Exception as e: print("$label failed", e)
PY
}
run base X=1
tail -3 $out/r2z_bench.err
