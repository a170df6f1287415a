#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python tools/trace_gemm.py --time-only > $out/r4d_time_gemm.txt 2>&1; echo "time exit $?"; cat $out/r4d_time_gemm.txt
timeout 300 python tools/trace_gemm.py l2.conv2 l3.conv2 l3.conv2.dgrad l3.conv2.wgrad l3.conv3.wgrad l2.conv1.wgrad enc.qk txt.qkv > $out/r4d_trace_gemm.txt 2>&1; echo "trace exit $?"; grep "==\|steady" $out/r4d_trace_gemm.txt
timeout 600 python tools/profile_step.py --out $out/r4d_timeline > $out/r4d_timeline.log 2>&1; echo "timeline exit $?"; head -14 $out/r4d_timeline.log | tail -10
