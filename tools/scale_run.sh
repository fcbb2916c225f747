#!/bin/bash
# 2/4/8-GPU bench lines + the reference arm on one 8-GPU box: gpurun --gpus 8 -- bash tools/scale_run.sh
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n=$n rc=$?"; tail -c 300 gpurun_out/scale_n$n.err | tail -2
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/scale_ref8.json 2>gpurun_out/scale_ref8.err; echo "ref rc=$?"
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/scale_n{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d.get("e2e_int8_llr",{}).get("value"), d["clocks"])
    except Exception as e: print(n,"fail",e)
print(open("gpurun_out/scale_ref8.json").read()[:400])
PY
