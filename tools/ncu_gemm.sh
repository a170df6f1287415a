#!/bin/bash
# Developer tool (GPU box): ncu --set full on a few standalone GEMM launches of one shape (fwd, dgrad, wgrad).
tag=${1:-l1.conv3}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:umma_gemm --launch-skip 2 --launch-count 5 \
    -o $out/ncu_${tag} -f python tools/prof_gemm.py --reps 1 --only $tag > $out/ncu_${tag}.log 2>&1
echo "ncu exit $?"
tail -3 $out/ncu_${tag}.log
