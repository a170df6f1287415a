#!/bin/bash
# N = 4 sanity run of the default bench line (every extra measurement that runs at N > 1)
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
GPV_BENCH_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 > $out/r3m_bench_n4.json 2> $out/r3m_bench_n4.err; echo "bench exit $?"
python - <<PY
import json
txt=open("$out/r3m_bench_n4.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
for k in ["value","ms_per_step","n_gpus","grads_equal_across_ranks","allreduce_bytes_per_step"]:
    print(k, json.dumps(d.get(k))[:300])
print("e2e", d["e2e"]["value"], "multitask", d["multitask"]["value"], d["multitask"]["answer_lengths"], "full", d["full_step"]["ms_per_step"])
PY
grep "bench rank 0" $out/r3m_bench_n4.err | tail -4
