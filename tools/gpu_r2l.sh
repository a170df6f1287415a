#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_layer_gpu.py -x -q -rP -m gpu -k attn > $out/r2o_attn.log 2>&1; echo "pytest exit $?"; grep "attn_block" $out/r2o_attn.log | head -20; tail -3 $out/r2o_attn.log
timeout 300 python tools/prof_layer.py > $out/r2o_prof_layer.txt 2>&1; grep attn_block $out/r2o_prof_layer.txt
timeout 120 python tools/trace_layer.py > $out/r2o_trace.txt 2>&1; tail -10 $out/r2o_trace.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_block_fwd_kernel -c 1 -o $out/r2o_attn_block --force-overwrite python tools/prof_layer.py > $out/r2o_ncu.log 2>&1; echo "ncu exit $?"; tail -3 $out/r2o_ncu.log
