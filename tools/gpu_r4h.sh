#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python tools/trace_gemm.py --time-only > $out/r4h_time_gemm.txt 2>&1; echo "time exit $?"; cat $out/r4h_time_gemm.txt
timeout 300 python tools/trace_gemm.py stem l3.conv2 l2.conv2 > $out/r4h_trace_gemm.txt 2>&1; echo "trace exit $?"; grep "==\|steady" $out/r4h_trace_gemm.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/r4h_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r4h_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4h_bench.json 2> $out/r4h_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r4h_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
