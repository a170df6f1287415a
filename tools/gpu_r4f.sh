#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python tools/trace_gemm.py stem l1.conv2 l3.conv2 l2.conv2 > $out/r4f_trace_gemm.txt 2>&1; echo "trace exit $?"; grep "==\|steady" $out/r4f_trace_gemm.txt
