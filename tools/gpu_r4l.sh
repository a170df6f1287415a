#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r4l_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r4l_pytest.log
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r4l_bench.json 2> $out/r4l_bench.err; echo "bench exit $?"; tail -3 $out/r4l_bench.err
python - <<PY
import json
d=json.load(open("$out/r4l_bench.json"))
for k in ["value","ms_per_step","e2e","e2e_sync_read","roofline","roofline_step","encdec_block","multitask","decode","full_step","gpu_launches","host_enqueue_ms_per_step"]:
    print(k, json.dumps(d.get(k))[:420])
PY
