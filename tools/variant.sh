#!/bin/bash
# build a variant of the library that differs in the flags of ONE source file (default: instances_scalar2d.cu, the
# config-2 kernels): tools/variant.sh TAG "-DFLAG=1 ..." [source.cu]  ->  mfem-ad_b200/libmadb_TAG.so
# (run a program on it with MADB_LIB=libmadb_TAG.so; the other objects come from the last full build)
set -e
cd "$(dirname "$0")/../mfem-ad_b200"
TAG=$1; FL=$2; SRC=${3:-instances_scalar2d.cu}
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC \
  -diag-suppress 177,550,128 $FL -x cu -c csrc/$SRC -o build/${SRC}.${TAG}.o
OBJS=$(ls build/*.cu.o build/*.cpp.o | grep -v "build/${SRC}.o")
nvcc -shared -o libmadb_${TAG}.so $OBJS build/${SRC}.${TAG}.o -gencode arch=compute_100a,code=sm_100a
echo built libmadb_${TAG}.so
