#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 75 ncu --set full --clock-control none -k regex:'attn_block_fwd_kernel|mlp_block_fwd_kernel' -o /tmp/r5g_layer -f python tools/ncu_layer_once.py > $out/r5g_ncu_layer.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r5g_layer.ncu-rep --page raw --csv > $out/r5g_ncu_layer_raw.csv 2>/dev/null; python tools/ncu_digest.py $out/r5g_ncu_layer_raw.csv > $out/r5g_ncu_layer_digest.txt 2>&1; grep -c kernel: $out/r5g_ncu_layer_digest.txt; tail -5 $out/r5g_ncu_layer.log | cut -c1-200
