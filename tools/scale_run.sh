#!/bin/bash
# BASELINE configs 4 and 5 as specified, on ONE 8-GPU box:   gpurun --gpus 8 -- bash tools/scale_run.sh
#   config 4: scaled min-sum BG2 Z=32, 10 it, Eb/N0 sweep, 10^9 frames split over 1/2/4/8 GPUs (tools/sweep_run.py; the
#             reduced counters must be IDENTICAL for every N -- Philox is keyed by the global frame index)
#   config 5: message-GNN training step (fwd + bwd + SGD, gradient all-reduce) at 1/2/4/8 GPUs (tools/train_bench.py)
#   plus the LDPCNeuralDecoder training step at 1/2/4/8 (QC-structured forward + backward kernels).
# Results: gpurun_out/scale_*.json; summary on stdout.
FRAMES=${FRAMES:-1e9}
PORT=29520
run() { n=$1; shift; if [ "$n" = 1 ]; then timeout 600 python "$@"; else PORT=$((PORT+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT "$@"; fi; }
for n in 1 2 4 8; do
  run $n tools/sweep_run.py --frames $FRAMES --out gpurun_out/scale_sweep_n$n.json > gpurun_out/scale_sweep_n$n.log 2> gpurun_out/scale_sweep_n$n.err; echo "sweep n=$n rc=$?"
done
for n in 1 2 4 8; do
  run $n tools/train_bench.py --model gnn --batch 512 --steps 5 > gpurun_out/scale_gnn_train_n$n.log 2> gpurun_out/scale_gnn_train_n$n.err; echo "gnn train n=$n rc=$?"
  run $n tools/train_bench.py --model neural --batch 32768 --steps 5 > gpurun_out/scale_neural_train_n$n.log 2> gpurun_out/scale_neural_train_n$n.err; echo "neural train n=$n rc=$?"
done
python - <<'PY'
import json, glob
sha = {}
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_sweep_n{n}.json"))
        sha[n] = d["counters_sha256"]
        print("sweep", n, "GPUs:", f"{d['frames_total']:.3g} frames {d['seconds']:.2f} s {d['codewords_per_s'] / 1e6:.1f} M cw/s {d['info_gbit_s']:.2f} Gbit/s", d["counters_sha256"][:12])
    except Exception as e:
        print("sweep", n, "failed", e)
print("counters identical for every N:", len(set(sha.values())) == 1 and len(sha) == 4)
for model in ("gnn", "neural"):
    for n in (1, 2, 4, 8):
        try:
            print(model, "train", n, open(f"gpurun_out/scale_{model}_train_n{n}.log").read().strip().splitlines()[-1][:300])
        except Exception as e:
            print(model, "train", n, "failed", e)
PY
