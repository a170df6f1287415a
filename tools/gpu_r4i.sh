#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4i_bench_$tag.json 2>> $out/r4i_bench.err; python - <<PY
import json
d=json.load(open("$out/r4i_bench_$tag.json"))
print("$tag", d["ms_per_step"], d["value"])
PY
}
run base X=1
run nobres GPVB200_BRES=0
run t1 GPVB200_TITER=0.20,0.0004
run t2 GPVB200_TITER=0.16,0.0006
run t3 GPVB200_TITER=0.20,0.0010
run base2 X=1
