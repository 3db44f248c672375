#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFT_REFILL_MIN=8 -DFT_MINB=8" [source.cu]   -> build/lib_NAME.so
# (only `source.cu`, default kernels_strict.cu, is recompiled with the extra defines; the other objects come from the in-tree build)
set -e
cd "$(dirname "$0")/../fraytracer_b200/csrc"
mkdir -p ../../build
NAME=$1; DEFS=$2; SRC=${3:-kernels_strict.cu}
BASE=$(basename $SRC .cu)
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall --expt-relaxed-constexpr \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false $DEFS -Xptxas -v -c $SRC -o ../../build/${BASE}_$NAME.o 2> ../../build/${BASE}_$NAME.ptxas.log
OBJS=$(ls *.o | grep -v "^$BASE.o")
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o ../../build/lib_$NAME.so $OBJS ../../build/${BASE}_$NAME.o
grep -A2 -E "Compiling entry function '_ZN(9ft_strict14machine_kernelILb0ELb1ELb1ENS_10RenderJobTILb0|8ft_cloud12cloud_kernelILb1ELb0|11ft_spectral21spectral_trace_kernelILb1ELb0ELb0)" ../../build/${BASE}_$NAME.ptxas.log | grep -E "registers|spill" | head -4
