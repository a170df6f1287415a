#!/bin/bash
# Developer tool (GPU box): ncu --set full over the non-GEMM kernels (one launch each at bench shape).
out=gpurun_out
mkdir -p $out
python tools/prof_kernels.py > $out/${1:-rX}_kernels.txt 2>&1
timeout 900 ncu --set full --clock-control none -k 'regex:attn_|layernorm|colsum|stem_s2d|maxpool|matcher|lsap|ce_fwd' -o $out/ncu_kernels -f python tools/prof_kernels.py --once > $out/ncu_kernels.log 2>&1
echo "ncu exit $?"
