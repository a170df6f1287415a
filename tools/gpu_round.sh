#!/bin/bash
# Developer tool (GPU box): one gpurun call = parity tests + bench + per-entry breakdown + ncu launch list.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
GPV_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"
cat $out/${tag}_bench.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown $out/${tag}_breakdown.json > $out/${tag}_bench_eager.json 2>> $out/${tag}_bench.err
echo "breakdown exit $?"
timeout 300 python tools/prof_gemm.py > $out/${tag}_gemm_shapes.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --profiling --no-graph > $out/${tag}_ncu_bench.log 2>&1
echo "ncu exit $?"
