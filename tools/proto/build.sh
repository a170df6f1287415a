#!/bin/bash
# Developer tool: build the standalone kernel prototypes (they link the production library for the side-by-side timing).
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/proto/_build
for f in tools/proto/*.cu; do
  b=$(basename "$f" .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "$f" -o tools/proto/_build/$b \
       -Lgpv-1_b200/lib -lgpvb200 -lcuda -Xlinker -rpath -Xlinker '$ORIGIN/../../../gpv-1_b200/lib'
done
