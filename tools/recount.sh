#!/bin/bash
# Regenerate the roofline count files bench.py reads from the ncu reports of the LAST capture run
# (gpurun_out/minsum_r2.csv, gpurun_out/bp_fast_r2.csv; tools/capture_counts.sh makes them on the GPU box).
cd "$(dirname "$0")/.."
C=ldpc-neuralnetwork-decoder_b200/csrc
SRC="$C/decode_fast_kernel.cuh $C/bg2_tables.h $C/fast_kernels.cu $C/channel.cuh $C/params.cuh"
python tools/ncu_to_profile.py gpurun_out/minsum_r2.csv --codewords 262144 --kernel 'decode_fast_kernel<BG2Z32, 8, 0, 0>' --sources $SRC \
  --out profiles/r2_minsum_counts.json --note "scaled min-sum, BG2 Z=32, 10 iterations, 262144 codewords (tools/kbench.py)"
python tools/ncu_to_profile.py gpurun_out/bp_fast_r2.csv --codewords 262144 --kernel 'decode_fast_kernel<BG2Z32, 8, 1, 0>' --sources $SRC \
  --out profiles/r2_bp_counts.json --note "sum-product, BG2 Z=32, 10 iterations, 262144 codewords (tools/kbench.py)"
