#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py tests/test_blocks_gpu.py tests/test_layer_gpu.py -x -q -m gpu > $out/r2y_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r2y_pytest.log
timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2y_bench.json 2>> $out/r2y_bench.err
python - <<PY
import json
d=json.load(open("$out/r2y_bench.json")); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), {k:round(v,4) if isinstance(v,float) else v for k,v in d["roofline"].items() if k in ("achieved","frac","launches_per_step","avg_launch_us")})
a=d["attention_kernel"]
for k,v in a.items():
    if isinstance(v,dict): print(k, round(v["avg_launch_us"],1), "us", round(v["achieved"],1), "TFLOP/s")
PY
