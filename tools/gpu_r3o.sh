#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py -x -q -m gpu > $out/r3o_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r3o_pytest.log
timeout 300 python tools/bench_decode.py > $out/r3o_decode.json 2>&1; cut -c1-200 $out/r3o_decode.json
