#!/bin/bash
# round-2 GPU call A: fused mlp_block kernel tests + timing, full GPU suite with achieved parity errors, short bench
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2a_smi.txt 2>&1
timeout 300 python -m pytest tests/test_layer_gpu.py -x -q -s > $out/r2a_layer.log 2>&1; echo "layer tests exit $?"; tail -15 $out/r2a_layer.log
timeout 200 python tools/prof_layer.py > $out/r2a_prof_layer.txt 2>&1; echo "prof exit $?"; cat $out/r2a_prof_layer.txt
timeout 1200 python -m pytest tests -m gpu -q -s > $out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -8 $out/r2a_pytest.log; grep "\[parity\]" $out/r2a_pytest.log | head -80
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2a_bench.json 2> $out/r2a_bench.err; echo "bench exit $?"; cat $out/r2a_bench.json
GPVB200_FUSED=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2a_bench_unfused.json 2>> $out/r2a_bench.err; cat $out/r2a_bench_unfused.json
