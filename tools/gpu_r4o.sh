#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_entrypoints.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 > $out/r4o_bench.json 2> $out/r4o_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r4o_bench.json"))
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync", d["e2e_sync_read"]["ms_per_step"], "u8 no prefetch", d["e2e_uint8"]["ms_per_step"])
PY
