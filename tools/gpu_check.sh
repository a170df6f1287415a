#!/bin/bash
# Developer tool (GPU box): short validation call = pair-variant check + parity tests + default bench line (+ multitask with MT=1).
# Usage: gpurun --timeout 420 -- 'bash tools/gpu_check.sh <tag>'
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 150 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; tail -3 $out/${tag}_bench.err; cat $out/${tag}_bench.json
if [ -n "$MT" ]; then
  timeout 120 python bench.py --workload multitask --steps 16 --warmup 8 --no-cpu-baseline > $out/${tag}_bench_multitask.json 2> $out/${tag}_multitask.err
  echo "multitask exit $?"; tail -3 $out/${tag}_multitask.err; cat $out/${tag}_bench_multitask.json
fi
