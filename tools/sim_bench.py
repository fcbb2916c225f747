#!/usr/bin/env python
"""Throughput of the fused Monte-Carlo path (ldpc_sim_fer: generate -> decode -> count)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200.sim import simulate_fer
from ldpc_b200.utils import QCCode
code = QCCode.nr_2_0(32)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
simulate_fer(code, [-2.0], 1 << 20, device="cuda:0")
torch.cuda.synchronize(); t0 = time.perf_counter()
r = simulate_fer(code, [-2.0], n, device="cuda:0")
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"frames": n, "seconds": dt, "Mcw_s": n / dt / 1e6, "fer": r[0]["fer"], "frame_errors": r[0]["frame_errors"]}))
