#!/bin/bash
# Developer tool (GPU box): one gpurun call = parity tests + bench + timeline + pair-variant experiments + ncu launch list.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh <tag>'     (about 6 minutes of box time)
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
# per-kernel timeline of the graph-replayed step (idle time, lane overlap, SM occupancy estimate)
timeout 300 python tools/profile_step.py --out $out/${tag}_timeline > $out/${tag}_timeline.log 2>&1
echo "timeline exit $?"
# CTA-pair variant: the validated set, then the experimental weight-gradient extension, then both inside the step
timeout 200 python tools/check_pair.py > $out/${tag}_pair.log 2>&1; echo "pair check exit $?"
# production kernel, single-CTA vs pair, shape by shape (graph-replayed launches)
timeout 300 python tools/prof_pair.py > $out/${tag}_prof_pair.txt 2>&1; cat $out/${tag}_prof_pair.txt
GPVB200_PAIR=16 GPVB200_PAIR_BN=128 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_pair16_bn128.json 2>> $out/${tag}_bench.err
for thr in 16 32; do
  GPVB200_PAIR=$thr timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_pair$thr.json 2>> $out/${tag}_bench.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob("$out/${tag}_bench*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["ms_per_step"], 3), "ms", round(d["value"], 1), "samples/s")
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown $out/${tag}_breakdown.json > $out/${tag}_bench_eager.json 2>> $out/${tag}_bench.err
echo "breakdown exit $?"
timeout 300 python tools/prof_gemm.py > $out/${tag}_gemm_shapes.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --profiling --no-graph > $out/${tag}_ncu_bench.log 2>&1
echo "ncu exit $?"
python tools/summarize_launches.py $out/${tag}_launches.csv --traffic $out/${tag}_step_traffic.json > $out/${tag}_launches.md 2>&1
