#!/bin/bash
# usage: tools/build_variant.sh NAME [-D...]   -> build_variants/libldpc_b200_NAME.so
set -e
cd "$(dirname "$0")/.."
N=$1; shift
P=ldpc-neuralnetwork-decoder_b200
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC"
mkdir -p build_variants/obj
if [ "${TU:-fast}" = "main" ]; then
  nvcc $F "$@" -c -o build_variants/obj/main_$N.o $P/csrc/ldpc_b200.cu
  nvcc -shared -o build_variants/libldpc_b200_$N.so build_variants/obj/main_$N.o $P/build/fast_kernels.o $P/build/neural_qc.o
else
  nvcc $F "$@" -c -o build_variants/obj/fast_$N.o $P/csrc/fast_kernels.cu
  [ -f $P/build/ldpc_b200.o ] || nvcc $F -c -o $P/build/ldpc_b200.o $P/csrc/ldpc_b200.cu
  nvcc -shared -o build_variants/libldpc_b200_$N.so $P/build/ldpc_b200.o build_variants/obj/fast_$N.o $P/build/neural_qc.o
fi
echo built $N
