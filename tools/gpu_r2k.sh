#!/bin/bash
# round-2 GPU call K: new decode kernels + tightened parity tests, bench line with the extra measurements
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py tests/test_blocks_gpu.py -x -q -rP -m gpu > $out/r2k_pytest.log 2>&1; echo "pytest exit $?"; grep "parity\]" $out/r2k_pytest.log | sort -u | tail -60; tail -15 $out/r2k_pytest.log
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r2k_bench.json 2> $out/r2k_bench.err; echo "bench exit $?"; tail -5 $out/r2k_bench.err
python - <<PY
import json
d=json.load(open("$out/r2k_bench.json"))
for k in ["value","ms_per_step","e2e","e2e_fp32_sync","multitask","decode","torch_eager_gpu","cpu_baseline"]:
    print(k, json.dumps(d.get(k))[:700])
PY
