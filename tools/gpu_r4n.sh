#!/bin/bash
# two GPUs: DDP sanity after the kernel / loop changes of this session (default bench line at N = 2, the 2-GPU NCCL gradient test)
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-extras --no-cpu-baseline > $out/r4n_bench_n2.json 2> $out/r4n_bench_n2.err; echo "bench N=2 exit $?"
python - <<PY
import json
d=json.load(open("$out/r4n_bench_n2.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("grads_equal_across_ranks"))
PY
timeout 300 python -m pytest tests/test_parallel.py -m gpu -x -q 2>&1 | tail -2
