#!/bin/bash
# two GPUs: the bench line of the closing build at N = 2 (per-step event marks in the timed region, gradient check across ranks)
out=gpurun_out; mkdir -p $out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $out/r5f_bench_n2.json 2> $out/r5f_bench_n2.err; echo "bench N=2 exit $?"
python - <<PY
import json
d=json.load(open("$out/r5f_bench_n2.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("grads_equal_across_ranks"), d.get("step_ms"))
PY
tail -3 $out/r5f_bench_n2.err
