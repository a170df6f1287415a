#!/bin/bash
out=gpurun_out; mkdir -p $out
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r3l_bench.json 2> $out/r3l_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r3l_bench.json"))
for k in ["value","ms_per_step","decode","encdec_block"]:
    print(k, json.dumps(d.get(k))[:700])
print("traffic", d["roofline"]["traffic"], "e2e", d["e2e"]["value"])
PY
