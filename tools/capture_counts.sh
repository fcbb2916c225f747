#!/bin/bash
# On the GPU box (gpurun -- bash tools/capture_counts.sh): one `ncu --set full` capture of the headline min-sum kernel and of
# the specialised sum-product kernel, each after the same command has exited 0 without ncu.  Then, back in the build
# container: bash tools/recount.sh
P=ldpc-neuralnetwork-decoder_b200/libldpc_b200.so
for algo in minsum bp; do
  python tools/kbench.py $P --algo $algo --batch 262144 --reps 2 || exit 1
  out=gpurun_out/$( [ $algo = minsum ] && echo minsum_r2 || echo bp_fast_r2 )
  ncu --set full --clock-control none -k regex:decode_fast_kernel -s 2 -c 1 -o $out -f python tools/kbench.py $P --algo $algo --batch 262144 --reps 1 > gpurun_out/ncu_$algo.log 2>&1
  # the reports of these 250 KB-SASS kernels are 40 MB each and gpurun brings back at most 64 MB: keep the raw-page CSV only
  ncu -i $out.ncu-rep --page raw --csv > $out.csv 2>/dev/null && rm -f $out.ncu-rep
done
