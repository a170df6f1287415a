#!/bin/bash
# round-2 GPU call P: attn_block in the encoder (tests), scheduling knobs inside the step
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_dropout_gpu.py tests/test_layer_gpu.py tests/test_model_gpu.py -x -q -m gpu > $out/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2p_pytest.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2p_bench_$label.json 2>> $out/r2p_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r2p_bench_$label.json")); print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), " launches", d["gpu_launches"])
except Exception as e: print("$label failed", e)
PY
}
run base X=1
run noattnblock GPVB200_ATTN_BLOCK=0
run kper4 GPVB200_MIN_KPER=4
run kper8 GPVB200_MIN_KPER=8
run wg96 GPVB200_WGRAD_CTAS=96
run wg64 GPVB200_WGRAD_CTAS=64
run prio GPVB200_MAIN_PRIO=1
run prio_kper4 GPVB200_MAIN_PRIO=1 GPVB200_MIN_KPER=4
run prio_wg96 GPVB200_MAIN_PRIO=1 GPVB200_WGRAD_CTAS=96
