#!/bin/bash
# compute-sanitizer, second half: memcheck over the block / end-to-end / matcher tests, racecheck (shared-memory hazards) over the
# kernels that do not synchronise through mbarriers (element-wise, norm, loss, matcher / assignment, decode)
out=gpurun_out; mkdir -p $out
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_blocks_gpu.py tests/test_model_gpu.py tests/test_matcher.py tests/test_entrypoints.py -m gpu -x -q > $out/r5d_memcheck_model.log 2>&1
echo "memcheck model exit $?"; tail -4 $out/r5d_memcheck_model.log | cut -c1-300
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_ops_gpu.py tests/test_matcher.py tests/test_decode_gpu.py -m gpu -x -q > $out/r5d_racecheck_ops.log 2>&1
echo "racecheck ops exit $?"; grep -m5 -A6 "hazard" $out/r5d_racecheck_ops.log | cut -c1-300; tail -4 $out/r5d_racecheck_ops.log | cut -c1-300
