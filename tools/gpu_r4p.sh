#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python tools/trace_gemm.py --time-only l1.conv2 stem l3.conv1 l2.conv2 l3.conv2 l3.conv2.dgrad l3.conv2.wgrad l3.conv3.wgrad l2.conv1.wgrad > $out/r4p_time_gemm.txt 2>&1; echo "time exit $?"; cat $out/r4p_time_gemm.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/r4p_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r4p_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4p_bench.json 2> $out/r4p_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r4p_bench.json"))
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
