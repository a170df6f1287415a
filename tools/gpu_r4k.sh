#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 60 tools/proto/_build/mma_issue > $out/r4j_mma_issue.txt 2>&1; cat $out/r4j_mma_issue.txt
S="l1.conv2 stem l3.conv1 l2.conv2 l3.conv2 l3.conv2.dgrad l3.conv2.wgrad"
for o in 0 1 2 3; do echo "MMAOPT=$o"; GPVB200_MMAOPT=$o timeout 300 python tools/trace_gemm.py --time-only $S 2>&1 | tee $out/r4k_time_opt$o.txt; done
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -2
GPVB200_MMAOPT=3 timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4k_bench_$tag.json 2>> $out/r4k_bench.err; python - <<PY
import json
d=json.load(open("$out/r4k_bench_$tag.json"))
print("$tag", d["ms_per_step"], d["value"])
PY
}
run base X=1
run opt1 GPVB200_MMAOPT=1
run opt3 GPVB200_MMAOPT=3
run nobres GPVB200_BRES=0
run t1 GPVB200_TITER=0.20,0.0004
run t2 GPVB200_TITER=0.16,0.0006
run base2 X=1
