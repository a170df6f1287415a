#!/bin/bash
# round-2 final pass (second session): full GPU suite, smoke, default bench line + reference arm, ncu launch list + DRAM traffic of THIS
# build, per-kernel timeline of the replayed step, ncu --set full of the deep trunk contraction
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r4m_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r4m_pytest.log
timeout 600 python __graft_entry__.py smoke > $out/r4m_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r4m_smoke.log
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r4m_bench.json 2> $out/r4m_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/r4m_bench_reference.json 2>> $out/r4m_bench.err; echo "reference arm exit $?"; cut -c1-300 $out/r4m_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $out/r4m_launches.csv \
    python bench.py --steps 1 --warmup 1 --profiling --no-graph > $out/r4m_ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py $out/r4m_launches.csv --traffic $out/r4m_step_traffic.json > $out/r4m_launches.md 2>&1; head -12 $out/r4m_launches.md; cat $out/r4m_step_traffic.json
timeout 600 python tools/profile_step.py --out $out/r4m_timeline > $out/r4m_timeline.log 2>&1; echo "timeline exit $?"; head -14 $out/r4m_timeline.log | tail -10
timeout 600 ncu --set full --import-source on --clock-control none -k regex:umma_gemm --launch-skip 3 --launch-count 2 -o $out/r4m_ncu_l3conv2 -f python tools/trace_gemm.py --time-only l3.conv2 > $out/r4m_ncu_l3conv2.log 2>&1; echo "ncu full exit $?"
ncu -i $out/r4m_ncu_l3conv2.ncu-rep --page raw --csv > $out/r4m_ncu_l3conv2_raw.csv 2>/dev/null; python tools/ncu_digest.py $out/r4m_ncu_l3conv2_raw.csv 2>&1 | head -40
python - <<PY
import json
d=json.load(open("$out/r4m_bench.json"))
for k in ["value","ms_per_step","e2e","e2e_sync_read","roofline","roofline_step","encdec_block","multitask","decode","torch_eager_gpu","cpu_baseline","full_step","gpu_launches","clocks"]:
    print(k, json.dumps(d.get(k))[:600])
PY
