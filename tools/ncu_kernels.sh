#!/bin/bash
# Developer tool (GPU box): ncu --set full over the non-GEMM kernels (one launch each at bench shape) and over two GEMM
# shapes; only the raw-page CSV exports travel back (the .ncu-rep files exceed gpurun's 64 MiB return limit).
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python tools/prof_kernels.py > $out/${tag}_kernels.txt 2>&1
timeout 900 ncu --set full --clock-control none -k 'regex:attn_|layernorm|colsum|stem_s2d|maxpool|matcher|lsap|ce_fwd' -o /tmp/ncu_kernels -f \
    python tools/prof_kernels.py --once > $out/ncu_kernels.log 2>&1
echo "ncu kernels exit $?"
ncu -i /tmp/ncu_kernels.ncu-rep --page raw --csv > $out/${tag}_kernels_ncu_raw.csv 2>/dev/null
for shape in l1.conv3 l3.conv1; do
  timeout 600 ncu --set full --clock-control none -k regex:umma_gemm --launch-skip 2 --launch-count 5 -o /tmp/ncu_$shape -f \
      python tools/prof_gemm.py --reps 1 --only $shape > $out/ncu_$shape.log 2>&1
  echo "ncu $shape exit $?"
  ncu -i /tmp/ncu_$shape.ncu-rep --page raw --csv > $out/${tag}_gemm_${shape}_ncu_raw.csv 2>/dev/null
done
ls -la $out
