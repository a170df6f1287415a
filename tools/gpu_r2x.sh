#!/bin/bash
# round-2 GPU call X: bench line with every extra measurement, ncu launch list + DRAM traffic of one step, ncu --set full of the
# dominant kernel on two trunk shapes and of the two fused layer kernels
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2x_smi.txt 2>&1
GPV_BENCH_VERBOSE=1 timeout 900 python bench.py > $out/r2x_bench.json 2> $out/r2x_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$out/r2x_bench.json"))
for k in ["value","ms_per_step","e2e","roofline","encdec_block","multitask","decode","torch_eager_gpu","cpu_baseline","attention_kernel"]:
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $out/r2x_launches.csv \
    python bench.py --steps 1 --warmup 1 --profiling --no-graph > $out/r2x_ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py $out/r2x_launches.csv --traffic $out/r2x_step_traffic.json > $out/r2x_launches.md 2>&1; head -30 $out/r2x_launches.md
for shape in l1.conv3 l3.conv1; do
  timeout 600 ncu --set full --clock-control none -k regex:umma_gemm --launch-skip 2 --launch-count 2 -o /tmp/ncu_$shape -f \
      python tools/prof_gemm.py --reps 1 --only $shape > $out/ncu_$shape.log 2>&1
  echo "ncu $shape exit $?"
  ncu -i /tmp/ncu_$shape.ncu-rep --page raw --csv > $out/r2x_gemm_${shape}_ncu_raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k regex:mlp_block_fwd_kernel -c 1 -o /tmp/ncu_mlp -f python tools/prof_layer.py > $out/ncu_mlp.log 2>&1
ncu -i /tmp/ncu_mlp.ncu-rep --page raw --csv > $out/r2x_mlp_block_ncu_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:attn_block_fwd_kernel -c 2 -o /tmp/ncu_attn -f python tools/prof_layer.py > $out/ncu_attn.log 2>&1
ncu -i /tmp/ncu_attn.ncu-rep --page raw --csv > $out/r2x_attn_block_ncu_raw.csv 2>/dev/null
ls -la $out | grep r2x
