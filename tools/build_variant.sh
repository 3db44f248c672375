#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFT_REFILL_MIN=8 -DFT_MINB=8"   -> build/lib_NAME.so (only kernels_strict.cu is recompiled)
set -e
cd "$(dirname "$0")/../fraytracer_b200/csrc"
mkdir -p ../../build
NAME=$1; DEFS=$2
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall --expt-relaxed-constexpr \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false $DEFS -Xptxas -v -c kernels_strict.cu -o ../../build/ks_$NAME.o 2> ../../build/ks_$NAME.ptxas.log
OBJS=$(ls *.o | grep -v kernels_strict.o)
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o ../../build/lib_$NAME.so $OBJS ../../build/ks_$NAME.o
grep -A2 "machine_kernelILb0ELb1ELb1ENS_10RenderJobTILb0" ../../build/ks_$NAME.ptxas.log | grep -E "registers|spill" | head -3
