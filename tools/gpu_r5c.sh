#!/bin/bash
# compute-sanitizer memcheck over the kernel-level GPU tests of the shipped library (out-of-bounds / misaligned global, shared and
# tensor-memory accesses in every hand-written kernel at the tests' shapes, ragged ones included)
out=gpurun_out; mkdir -p $out
san() { label=$1; shift
  timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest "$@" -m gpu -x -q > $out/r5c_memcheck_$label.log 2>&1
  echo "memcheck $label exit $?"; grep -c "Invalid\|Misaligned" $out/r5c_memcheck_$label.log; tail -4 $out/r5c_memcheck_$label.log | cut -c1-300
}
san ops tests/test_ops_gpu.py tests/test_decode_gpu.py tests/test_dropout_gpu.py
san layer tests/test_layer_gpu.py
san gemm tests/test_gemm_gpu.py
