#!/bin/bash
# args: M N K stages reps bar_mode flags kblk max_pairs
cd "$(dirname "$0")/_build"
run() { timeout 20 ./gemm_2cta "$@" | grep -v "^production   mismatches"; }
echo "== M sweep (full waves: 18944 = 74 pairs x 256)"
for M in 18944 37888 38400 56832; do run $M 256 1024 6 20 0 0 1; done
echo "== dissection at M=37888: flags 1 no MMA, 2 no TMA, 3 neither, 6 no TMA no stores, 4 no stores"
for f in 1 2 3 6 4; do run 37888 256 1024 6 20 0 $f 1; done
echo "== stage shape at M=37888"
run 37888 256 1024 4 20 0 0 1
run 37888 256 1024 3 20 0 0 2
run 37888 256 1024 2 20 0 0 3
echo "== half the chip (37 pairs), M=9472 and 18944"
run 9472 256 1024 6 20 0 0 1 37
run 18944 256 1024 6 20 0 0 1 37
echo "== deep K: M=18944 K=4096"
run 18944 256 4096 6 20 0 0 1
run 18944 256 4096 6 20 0 1 1
run 18944 256 4096 6 20 0 2 1
