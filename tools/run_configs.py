#!/usr/bin/env python
"""Run the five BASELINE.json configurations on one B200 and write profiles/r1_configs.json.

These are parity/acceptance cases, not bench lines (bench.py measures the headline metric).
FER-vs-reference: the reference's Python decoders manage ~5 codewords/s, so the engine's FER
curve is compared with the CPU oracle (pinned bit-for-bit to the reference) on a bounded sample
of the SAME frames (same Philox keying) -- exact counter equality up to the generator's libm/CUDA
last-ulp differences -- and with Wilson intervals on the large sample."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder, create_message_gnn_decoder
from ldpc_b200.sim import simulate_fer, wilson_interval
from ldpc_b200.training import train_step
from ldpc_b200.utils import QCCode
from oracle import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--frames4", type=float, default=1e9, help="total frames of config 4 (all SNR points)")
ap.add_argument("--frames2", type=int, default=1 << 20)
ap.add_argument("--gnn-batch", type=int, default=1 << 15)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_configs.json"))
a = ap.parse_args()
dev = torch.device("cuda", 0)
res = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t0


def oracle_counts(code, Z, algo, iters, snr_db, seed, frames, order):
    llr = oracle.awgn_llr(None, frames, code.N, snr_db, seed)
    o = oracle.decode(code.shifts, Z, llr, iters, algo, 0.75, order=order, threads=os.cpu_count())
    nerr = o["hard"].sum(axis=1)
    return int(nerr.sum()), int((nerr > 0).sum())

# ---- config 1: min-sum Z=4, B=1024, 5 it (plumbing + parity with the reference's own output) ----
g = np.load(os.path.join(ROOT, "tests", "golden", "classic_z4_b1024_it5.npz"))
code4 = QCCode.nr_2_0(4)
dec = MinSumScaledDecoder(code4, 5, 0.75, early_stopping=False)
llr = torch.from_numpy(g["llr"]).to(dev)
dec.decode(llr)
(bits, it), dt = timed(lambda: dec.decode(llr))
ref_bits = np.unpackbits(g["ms_bits"], axis=1)[:, :code4.N]
res["config1_minsum_z4_b1024_it5"] = {"decode_s": dt, "codewords_per_s": 1024 / dt, "reference_decode_s": float(g["ref_seconds"][1]),
                                      "hard_bit_mismatches_vs_reference": int((bits.cpu().numpy().astype(np.uint8) != ref_bits).sum())}

# ---- config 2: BP Z=32, 10 it, Eb/N0 0..4 dB, 2^20 codewords per point ----
code = QCCode.nr_2_0(32)
ebn0 = [0.0, 1.0, 2.0, 3.0, 4.0]
snrs = [e - 4.150 for e in ebn0]
pts, dt = timed(lambda: simulate_fer(code, snrs, a.frames2, algo="bp", iters=10, seed=1234, device=dev))
chk = []
for k, s in enumerate(snrs[:3]):
    be, fe = oracle_counts(code, 32, "bp", 10, s, 1234 + k, 4096, "reference")
    sub = simulate_fer(code, [s], 4096, algo="bp", iters=10, seed=1234 + k, device=dev)[0]
    chk.append({"snr_db": s, "oracle_frame_errors": fe, "engine_frame_errors": sub["frame_errors"],
                "oracle_bit_errors": be, "engine_bit_errors": sub["bit_errors"]})
res["config2_bp_z32_it10_sweep"] = {"ebn0_db": ebn0, "points": pts, "seconds": dt, "codewords_per_s": a.frames2 * len(snrs) / dt,
                                    "same_frames_vs_oracle_4096": chk}

# ---- config 3: GNN inference, 5 layers, hidden 64, 32 types, random init (seed 0) ----
torch.manual_seed(0)
gnn, conv = create_message_gnn_decoder(code, 5, 64, base_graph=code.base_matrix(), Z=32)
B3 = a.gnn_batch
llr3 = torch.empty((B3, code.N), dtype=torch.float32, device=dev)
_native.check(_native.lib().ldpc_awgn_llr(None, B3, code.N, -2.0, 7, 0, _native.ptr(llr3), None))
gnn(llr3[:256])
(out3, dt) = timed(lambda: gnn(llr3))
sd = {k: v.detach().numpy() for k, v in gnn.state_dict().items()}
soft_ref, _ = oracle.gnn_forward(sd, llr3[:8].cpu().numpy(), conv.message_var_index.numpy(), conv.message_check_index.numpy(),
                                 gnn._expanded_types(), 5)
rel = np.abs(out3[0][:8].cpu().numpy() - soft_ref) / np.maximum(np.abs(soft_ref), 1.0)
res["config3_gnn_inference"] = {"batch": B3, "seconds": dt, "codewords_per_s": B3 / dt, "seconds_for_2^20_extrapolated": dt * (1 << 20) / B3,
                                "max_rel_err_soft_vs_oracle_8cw": float(rel.max())}

# ---- config 4: min-sum Z=32, 10 it, FER sweep, frames4 frames in total ----
ebn0_4 = [1.0, 1.5, 2.0, 2.5, 3.0, 3.5]
snrs4 = [e - 4.150 for e in ebn0_4]
per_point = int(a.frames4 // len(snrs4))
pts4, dt4 = timed(lambda: simulate_fer(code, snrs4, per_point, algo="minsum", iters=10, seed=99, device=dev))
chk4 = []
for k, s in enumerate(snrs4[:4]):
    n = 20000
    be, fe = oracle_counts(code, 32, "minsum", 10, s, 99 + k, n, "reference")
    sub = simulate_fer(code, [s], n, algo="minsum", iters=10, seed=99 + k, device=dev)[0]
    lo, hi = wilson_interval(pts4[k]["frame_errors"], pts4[k]["frames"])
    chk4.append({"snr_db": s, "frames": n, "oracle_frame_errors": fe, "engine_frame_errors": sub["frame_errors"],
                 "oracle_bit_errors": be, "engine_bit_errors": sub["bit_errors"],
                 "oracle_fer": fe / n, "oracle_fer_ci": wilson_interval(fe, n), "engine_full_fer": pts4[k]["fer"], "engine_full_fer_ci": [lo, hi]})
res["config4_minsum_z32_it10_fer_sweep"] = {"ebn0_db": ebn0_4, "frames_per_point": per_point, "points": pts4, "seconds": dt4,
                                            "codewords_per_s": per_point * len(snrs4) / dt4,
                                            "info_gbit_s": per_point * len(snrs4) * 320 / dt4 / 1e9,
                                            "same_frames_vs_oracle_20000": chk4}

# ---- config 5: GNN training step (fwd + bwd + SGD), final-output mean BCE ----
Bt = 1024
gnn = gnn.cuda()
opt = torch.optim.SGD(gnn.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
gt = torch.zeros((Bt, code.N), device=dev)
l0 = float(train_step(gnn, llr3[:Bt], gt, opt))
(l1, dt5) = timed(lambda: float(train_step(gnn, llr3[:Bt], gt, opt)))
res["config5_gnn_training_step"] = {"batch": Bt, "seconds_per_step": dt5, "codewords_per_s": Bt / dt5, "loss_step0": l0, "loss_step1": l1,
                                    "grad_elements": sum(p.numel() for p in gnn.parameters())}
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk not in ("points",)}) for k, v in res.items()}, indent=1)[:6000])
