#!/bin/bash
# two GPUs of one box: the NCCL gradient-equality test that is skipped on one GPU, and the bench line at N = 2
out=gpurun_out; mkdir -p $out
nvidia-smi -L
echo skip parallel test
GPV_BENCH_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $out/r3d_bench_n2.json 2> $out/r3d_bench_n2.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r3d_bench_n2.json"))
for k in ["value","ms_per_step","n_gpus","e2e","grads_equal_across_ranks","allreduce_bytes_per_step","multitask","full_step"]:
    print(k, json.dumps(d.get(k))[:300])
PY
grep "bench rank" $out/r3d_bench_n2.err | tail -30
