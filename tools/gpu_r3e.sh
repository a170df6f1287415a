#!/bin/bash
# N = 2: bucket merging of the gradient all-reduce (fewer graph boundaries in the captured backward)
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $out/r3e_bench_n2_$label.json 2> $out/r3e_bench_n2_$label.err; echo "$label exit $?"
  python - <<PY
import json
try:
    txt=open("$out/r3e_bench_n2_$label.json").read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "full", round(d["full_step"]["ms_per_step"],3), "grads_equal", d.get("grads_equal_across_ranks"))
except Exception as e: print("$label failed", e)
PY
}
run all7 X=1
run at3456 GPVB200_DDP_REDUCE_AT=3,4,5,6
run at13456 GPVB200_DDP_REDUCE_AT=1,3,4,5,6
run at356 GPVB200_DDP_REDUCE_AT=3,5,6
