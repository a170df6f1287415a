#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python tools/trace_gemm.py > $out/r4a_trace_gemm.txt 2>&1; echo "trace exit $?"; grep "==\|steady" $out/r4a_trace_gemm.txt
