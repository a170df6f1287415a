#!/bin/bash
# Fixed cost of one launch (prologue + one tile + drain), with and without programmatic dependent launch on the production kernel.
cd "$(dirname "$0")/_build"
run() { timeout 20 ./gemm_2cta "$@" | grep -v "mismatches"; }
for pdl in 1 0; do
  echo "== GPVB200_PDL=$pdl (production kernel only; the prototype launches without PDL)"
  export GPVB200_PDL=$pdl
  echo "-- one pair tile, one stage: M=256 N=256 K=64"; run 256 256 64 3 50 0 0 1
  echo "-- one full wave, one stage: M=18944 N=256 K=64"; run 18944 256 64 3 50 0 0 1
  echo "-- one full wave, K=256: M=18944 N=256 K=256"; run 18944 256 256 3 50 0 0 2
  echo "-- small M like the decoders: M=3200 N=256 K=256 / M=640 N=768 K=768"; run 3200 256 256 3 50 0 0 2; run 640 768 768 3 50 0 0 2
done
