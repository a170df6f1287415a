#!/bin/bash
# N = 2 after the lane-join change: NCCL gradient-equality test + bench line
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_parallel.py -m gpu -x -q > $out/r3r_pytest_parallel.log 2>&1; echo "pytest exit $?"; tail -2 $out/r3r_pytest_parallel.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $out/r3r_bench_n2.json 2> $out/r3r_bench_n2.err; echo "bench exit $?"
python - <<PY
import json
txt=open("$out/r3r_bench_n2.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "full", round(d["full_step"]["ms_per_step"],3), "grads_equal", d.get("grads_equal_across_ranks"))
PY
