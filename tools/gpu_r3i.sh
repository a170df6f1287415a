#!/bin/bash
# ncu source-level profile of a narrow (N = 64) tile GEMM: layer1 conv1 as a plain GEMM, M = 614400, N = 64, K = 256
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:umma_gemm --launch-skip 2 --launch-count 1 -o $out/r3i_l1conv1 -f \
    python tools/prof_gemm.py --reps 1 --only l1.conv1 > $out/r3i_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $out/r3i_ncu.log
ncu -i $out/r3i_l1conv1.ncu-rep --page raw --csv > $out/r3i_l1conv1_raw.csv 2>/dev/null
ncu -i $out/r3i_l1conv1.ncu-rep --page source --csv > $out/r3i_l1conv1_source.csv 2>/dev/null
rm -f $out/r3i_l1conv1.ncu-rep
ls -la $out | grep r3i
