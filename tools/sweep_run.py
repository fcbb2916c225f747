#!/usr/bin/env python
"""BASELINE config 4 as specified: scaled min-sum BG2 Z=32, 10 iterations, Eb/N0 sweep, `--frames` all-zero codewords
in total, sharded over the ranks of one box with ONE all-reduce of the int64 error counters per block.

    python tools/sweep_run.py --frames 1e9                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           tools/sweep_run.py --frames 1e9 --out gpurun_out/sweep_nN.json

Runs sim.simulate_fer (fused generate -> decode -> count kernel, Philox keyed by the GLOBAL frame index) and writes
the per-point counters, the device-timed rate and the counters' SHA so that runs on 1/2/4/8 GPUs can be compared for
exact equality.  --checkpoint makes the sweep resumable (sim.py)."""
import argparse, hashlib, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200.sim import simulate_fer
from ldpc_b200.utils import QCCode

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=float, default=1e9, help="total frames over all SNR points")
ap.add_argument("--ebn0", type=float, nargs="*", default=[1.0, 1.5, 2.0, 2.5, 3.0, 3.5])
ap.add_argument("--algo", default="minsum")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--seed", type=int, default=99)
ap.add_argument("--checkpoint", default=None)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
code = QCCode.nr_2_0(32)
snrs = [e - 4.150 for e in a.ebn0]                  # snr_db = Eb/N0 + 10 log10(R), R = 10/26 per SURVEY 8d (-4.150 dB)
per_point = int(a.frames // len(snrs))
simulate_fer(code, snrs[:1], 1 << 16, algo=a.algo, iters=a.iters, seed=1, device=dev, rank=rank, world=world)     # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
pts = simulate_fer(code, snrs, per_point, algo=a.algo, iters=a.iters, seed=a.seed, device=dev, rank=rank, world=world,
                   checkpoint=a.checkpoint)
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
if rank == 0:
    counters = [[p["bit_errors"], p["frame_errors"], p["frames"], p["undetected"]] for p in pts]
    res = {"n_gpus": world, "algo": a.algo, "iters": a.iters, "ebn0_db": a.ebn0, "frames_per_point": per_point,
           "frames_total": per_point * len(snrs), "seconds": dt, "codewords_per_s": per_point * len(snrs) / dt,
           "info_gbit_s": per_point * len(snrs) * 320 / dt / 1e9, "points": pts,
           "counters_sha256": hashlib.sha256(json.dumps(counters).encode()).hexdigest()}
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "points"}))
    print(json.dumps([{k: p[k] for k in ("snr_db", "frames", "frame_errors", "bit_errors", "fer")} for p in pts]))
if world > 1:
    dist.destroy_process_group()
