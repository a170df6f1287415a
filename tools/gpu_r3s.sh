#!/bin/bash
# round-2 final pass: full GPU suite, smoke, default bench line, ncu launch list + DRAM traffic of THIS build
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r3s_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3s_pytest.log
timeout 600 python __graft_entry__.py smoke > $out/r3s_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r3s_smoke.log
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r3s_bench.json 2> $out/r3s_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/r3s_bench_reference.json 2>> $out/r3s_bench.err; echo "reference arm exit $?"; cut -c1-400 $out/r3s_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $out/r3s_launches.csv \
    python bench.py --steps 1 --warmup 1 --profiling --no-graph > $out/r3s_ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py $out/r3s_launches.csv --traffic $out/r3s_step_traffic.json > $out/r3s_launches.md 2>&1; head -12 $out/r3s_launches.md; cat $out/r3s_step_traffic.json
python - <<PY
import json
d=json.load(open("$out/r3s_bench.json"))
for k in ["value","ms_per_step","e2e","roofline","roofline_step","encdec_block","multitask","decode","torch_eager_gpu","cpu_baseline","full_step","gpu_launches","clocks"]:
    print(k, json.dumps(d.get(k))[:600])
PY
