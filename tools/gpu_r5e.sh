#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_entrypoints.py tests/test_model_gpu.py -m gpu -x -q > $out/r5e_pytest.log 2>&1; echo "pytest exit $?"; tail -15 $out/r5e_pytest.log | cut -c1-400
timeout 200 python tools/time_train_loop.py 24 > $out/r5e_train_loop.txt 2>&1; echo "train loop exit $?"; tail -4 $out/r5e_train_loop.txt | cut -c1-400
