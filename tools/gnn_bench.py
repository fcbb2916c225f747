#!/usr/bin/env python
"""GNN forward throughput (BASELINE config 3 shape: BG2 Z=32, 5 layers, hidden 64, 32 types)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
if os.environ.get("LDPC_LIB"):
    from ldpc_b200 import _native as _n
    _n.LIB_PATH = os.environ["LDPC_LIB"]
from ldpc_b200 import _native
from ldpc_b200.models import create_message_gnn_decoder
from ldpc_b200.utils import QCCode
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=2048); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--train", type=int, default=0)
a = ap.parse_args()
code = QCCode.nr_2_0(32)
torch.manual_seed(0)
dec, conv = create_message_gnn_decoder(code, 5, 64, base_graph=code.base_matrix(), Z=32)
dev = torch.device("cuda", 0)
B = a.batch
llr = torch.empty((B, code.N), dtype=torch.float32, device=dev)
_native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, -2.0, 1, 0, _native.ptr(llr), None))
dec(llr); torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec(llr); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
flop = 6304 * 5 * 12 * 64 * 64 * B          # E*L*12h^2 per codeword (MLPs as written)
print(json.dumps({"gnn_forward_ms": ms, "codewords_per_s": B / ms * 1e3, "TFLOPs_algorithmic": flop / ms / 1e9,
                  "frac_of_fp32_fma_peak_74.4": flop / ms / 1e9 / 74.4, "batch": B}))

if a.train:
    from ldpc_b200.training import train_step
    Bt = a.train
    dec = dec.cuda()
    opt = torch.optim.SGD(dec.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    gt = torch.zeros((Bt, code.N), device=dev)
    train_step(dec, llr[:Bt], gt, opt); torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = train_step(dec, llr[:Bt], gt, opt); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"gnn_train_step_ms": ms, "codewords_per_s": Bt / ms * 1e3, "batch": Bt, "loss": float(loss),
                      "TFLOPs_algorithmic_3x_fwd": 3 * 6304 * 5 * 12 * 64 * 64 * Bt / ms / 1e9}))
