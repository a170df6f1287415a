#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (tcgen05.mma = UTCHMMA, TMA = UTMALDG,
# tensor-memory load / store = LDTM / STTM, warp-level mma.sync = HMMA) in the shipped library.  Runs anywhere (cuobjdump only).
so=${1:-gpv-1_b200/lib/libgpvb200.so}
cuobjdump -sass "$so" | awk '
/Function : / { fn=$3; sub(/^_ZN3gpv[0-9]*/, "", fn); names[fn]=1; next }
/UTCHMMA/ { a[fn]++ } /UTMALDG/ { b[fn]++ } /LDTM/ { c[fn]++ } /STTM/ { d[fn]++ } /HMMA/ { if ($0 !~ /UTCHMMA/) e[fn]++ } /MUFU.EX2/ { f[fn]++ } /FFMA2|FADD2/ { g[fn]++ }
END { printf "%-90s %8s %8s %6s %6s %6s %9s %6s\n", "kernel (mangled, namespace stripped)", "UTCHMMA", "UTMALDG", "LDTM", "STTM", "HMMA", "MUFU.EX2", "F*2";
      for (k in names) printf "%-90s %8d %8d %6d %6d %6d %9d %6d\n", substr(k,1,90), a[k], b[k], c[k], d[k], e[k], f[k], g[k] }' | (read -r h; echo "$h"; sort)
