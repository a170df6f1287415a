#!/bin/bash
# round-2 GPU call J: full GPU suite, default bench, per-kernel timeline of the graph-replayed step
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -rP > $out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r2j_pytest.log
timeout 600 python bench.py > $out/r2j_bench.json 2> $out/r2j_bench.err; echo "bench exit $?"; cut -c1-600 $out/r2j_bench.json
timeout 600 python tools/profile_step.py --out $out/r2j_timeline > $out/r2j_timeline.log 2>&1; echo "timeline exit $?"; head -12 $out/r2j_timeline.log
