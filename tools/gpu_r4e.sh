#!/bin/bash
out=gpurun_out; mkdir -p $out
S="l1.conv1.k256 l1.conv1.k64 l1.conv2 stem l1.conv3 l2.conv2"
timeout 300 python tools/trace_gemm.py --time-only $S > $out/r4e_time_gemm.txt 2>&1; echo "time exit $?"; cat $out/r4e_time_gemm.txt
GPVB200_BRES=0 timeout 300 python tools/trace_gemm.py --time-only $S > $out/r4e_time_gemm_nobres.txt 2>&1; echo "time(no resident B) exit $?"; cat $out/r4e_time_gemm_nobres.txt
timeout 300 python tools/trace_gemm.py stem l1.conv2 l1.conv1.k64 > $out/r4e_trace_gemm.txt 2>&1; echo "trace exit $?"; grep "==\|steady" $out/r4e_trace_gemm.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/r4e_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r4e_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4e_bench.json 2> $out/r4e_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r4e_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
