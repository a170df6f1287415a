#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/r2s_pytest.log
