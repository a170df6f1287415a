#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 60 tools/proto/_build/mma_issue > $out/r4j_mma_issue.txt 2>&1; cat $out/r4j_mma_issue.txt
bash tools/gpu_r4i.sh
