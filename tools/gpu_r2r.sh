#!/bin/bash
# round-2 GPU call R: full GPU suite with the new defaults, bench (no extras), timeline
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2r_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2r_pytest.log
timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-extras > $out/r2r_bench.json 2> $out/r2r_bench.err; python - <<PY
import json
d=json.load(open("$out/r2r_bench.json")); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), "samples/s  e2e", round(d["e2e"]["value"],1), "full", round(d["full_step"]["ms_per_step"],3), {k:v for k,v in d["roofline"].items() if k in ("achieved","frac","launches_per_step","avg_launch_us")})
PY
timeout 600 python tools/profile_step.py --out $out/r2r_timeline > $out/r2r_timeline.log 2>&1; echo "timeline exit $?"
