#!/bin/bash
# The step's L2-/tensor-bound contraction shapes as plain GEMMs: pair kernel (3 stages x 64 KB) beside the production kernel.
# args: M N K stages reps bar_mode flags kblk max_pairs
cd "$(dirname "$0")/_build"
run() { timeout 20 ./gemm_2cta "$@" | grep -v "^production   mismatches"; }
echo "== layer3 conv2 3x3 (M=38400 N=256 K=2304), conv1 (K=1024), conv3 (N=1024 K=256)"
run 38400 256 2304 3 10 0 0 2
run 38400 256 1024 3 10 0 0 2
run 38400 1024 256 3 10 0 0 2
echo "== layer4 conv2 3x3 (M=9600 N=512 K=4608), conv1 (N=512 K=2048), conv3 (N=2048 K=512)"
run 9600 512 4608 3 10 0 0 2
run 9600 512 2048 3 10 0 0 2
run 9600 2048 512 3 10 0 0 2
echo "== encoder FFN (M=9600): ffn1 N=2048 K=256, ffn2 N=256 K=2048"
run 9600 2048 256 3 10 0 0 2
run 9600 256 2048 3 10 0 0 2
echo "== dissection with 64 KB stages at M=37888 K=1024: TMA only, MMA only, no stores"
run 37888 256 1024 3 10 0 1 2
run 37888 256 1024 3 10 0 2 2
run 37888 256 1024 3 10 0 4 2
