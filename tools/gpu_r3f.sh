#!/bin/bash
# N = 2: NCCL channel count (SMs the all-reduce kernels hold beside the backward pass)
out=gpurun_out; mkdir -p $out
run() { label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $out/r3f_bench_n2_$label.json 2> $out/r3f_bench_n2_$label.err; echo "$label exit $?"
  python - <<PY
import json
try:
    txt=open("$out/r3f_bench_n2_$label.json").read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print("$label", round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "full", round(d["full_step"]["ms_per_step"],3), "grads_equal", d.get("grads_equal_across_ranks"))
except Exception as e: print("$label failed", e)
PY
}
run default X=1
run ch4 NCCL_MAX_NCHANNELS=4
run ch8 NCCL_MAX_NCHANNELS=8
run ch16 NCCL_MAX_NCHANNELS=16
run ch8_cta NCCL_MAX_NCHANNELS=8 NCCL_NTHREADS=256
