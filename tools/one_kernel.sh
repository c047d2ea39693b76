#!/bin/bash
# quick look at the headline kernel alone: registers / spills (ptxas -v) and FP64 instruction mix (SASS) of
# k_patch_ws<MinimalSurfaceEnergy<2>, Q2> without rebuilding the library
cd "$(dirname "$0")/../mfem-ad_b200/build" || exit 1
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -diag-suppress 177,550,128 \
  -Xptxas -v $ONE_FLAGS -c one.cu -o one.o 2> one.log || { tail -20 one.log; exit 1; }
grep -A2 "k_patch_wsINS_20Minimal" one.log | grep -E "registers|spill"
cuobjdump -sass one.o | awk '/Function :/ {on = index($3, "k_patch_wsINS_20Minimal") > 0} on && /^ +\/\*[0-9a-f]+\*\// {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); c[op]++; t++} END {printf "total %d  DFMA %d DMUL %d DADD %d  LDS %d STS %d LDG %d STG %d MUFU %d\n", t, c["DFMA"], c["DMUL"], c["DADD"], c["LDS"], c["STS"], c["LDG"], c["STG"], c["MUFU"]}'
