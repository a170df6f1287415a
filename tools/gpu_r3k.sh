#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_decode_gpu.py -x -q -m gpu > $out/r3k_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/r3k_pytest.log
timeout 300 python tools/bench_decode.py > $out/r3k_decode.json 2>&1; cat $out/r3k_decode.json | cut -c1-260
