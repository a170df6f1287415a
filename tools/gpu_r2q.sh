#!/bin/bash
# round-2 GPU call Q: in-place stride-2 down-sample gradient (tests), lane / join scheduling knobs, pair variant in forward only
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_entrypoints.py -x -q -m gpu > $out/r2q_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2q_pytest.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2q_bench_$label.json 2>> $out/r2q_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2q_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), " loss", d["config"]["loss"])
except Exception as e: print("$label failed", e)
PY
}
run base X=1
run lazy GPVB200_LAZY_JOIN=1
run lanes2 GPVB200_WGRAD_LANES=2
run lazy_lanes2 GPVB200_LAZY_JOIN=1 GPVB200_WGRAD_LANES=2
run lazy_lanes3 GPVB200_LAZY_JOIN=1 GPVB200_WGRAD_LANES=3
run pairfwd16 GPVB200_PAIR=16 GPVB200_PAIR_FWD=1
run pairfwd8 GPVB200_PAIR=8 GPVB200_PAIR_FWD=1
tail -5 $out/r2q_bench.err
