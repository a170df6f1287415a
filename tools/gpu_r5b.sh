#!/bin/bash
# round-2 closing pass: full GPU suite, smoke, default bench line (+ step_ms spread, roofline.traffic attached by source hash) + reference
# arm, ncu --set full of a split-K weight-gradient launch and of the attention backward (the kernels DESIGN.md 8 names next)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r5b_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r5b_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/r5b_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/r5b_smoke.log
GPV_BENCH_VERBOSE=1 timeout 600 python bench.py > $out/r5b_bench.json 2> $out/r5b_bench.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $out/r5b_bench_reference.json 2>> $out/r5b_bench.err; echo "reference arm exit $?"; cut -c1-300 $out/r5b_bench_reference.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:umma_gemm --launch-skip 3 --launch-count 2 -o /tmp/r5b_wgrad -f python tools/trace_gemm.py --time-only l3.conv2.wgrad > $out/r5b_ncu_wgrad.log 2>&1; echo "ncu wgrad exit $?"
ncu -i /tmp/r5b_wgrad.ncu-rep --page raw --csv > $out/r5b_ncu_l3conv2_wgrad_raw.csv 2>/dev/null; python tools/ncu_digest.py $out/r5b_ncu_l3conv2_wgrad_raw.csv > $out/r5b_ncu_digest.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd -o /tmp/r5b_attn_bwd -f python tools/prof_kernels.py --once > $out/r5b_ncu_attn_bwd.log 2>&1; echo "ncu attn_bwd exit $?"
ncu -i /tmp/r5b_attn_bwd.ncu-rep --page raw --csv > $out/r5b_ncu_attn_bwd_raw.csv 2>/dev/null; python tools/ncu_digest.py $out/r5b_ncu_attn_bwd_raw.csv >> $out/r5b_ncu_digest.txt 2>&1
head -80 $out/r5b_ncu_digest.txt
python - <<PY
import json
d=json.load(open("$out/r5b_bench.json"))
for k in ["value","ms_per_step","step_ms","e2e","roofline","roofline_step","multitask","decode","full_step","gpu_launches","clocks"]:
    print(k, json.dumps(d.get(k))[:500])
PY
