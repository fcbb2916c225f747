#!/usr/bin/env python
"""Throughput of LDPCNeuralDecoder (unrolled neural min-sum, edge space) on one GPU:
fused variable+residual kernel vs the literal two-layer composition, inference and one
training step (forward + loss.mean().backward()).  BG2 Z=32 (E = 6304) or Z=4 (E = 788).
Usage: python tools/neural_bench.py [--batch 4096] [--iters 5] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ldpc_b200  # noqa: E402,F401
from ldpc_b200.models import LDPCNeuralDecoder  # noqa: E402
from ldpc_b200.utils import QCCode, create_LLR_mapping  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--z", type=int, default=32, choices=[4, 8, 16, 32], help="lifting size of the BG2 code")
    ap.add_argument("--lib", default=None, help="alternative libldpc_b200.so (tools/build_variant.sh)")
    args = ap.parse_args()
    if args.lib:
        from ldpc_b200 import _native
        _native.LIB_PATH = os.path.abspath(args.lib)
    dev = "cuda:0"
    code = QCCode.nr_2_0(args.z)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(dev), vidx.to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    llr = (torch.randn(args.batch, code.N, device=dev, generator=g) * 0.5 + 0.4)
    llr_e = llr[:, oidx[0].to(dev)].contiguous()
    gt = torch.ones_like(llr_e)
    out = {"batch": args.batch, "iters": args.iters, "Z": args.z, "E": code.E, "lib": args.lib}
    gt_v = torch.ones_like(llr)
    # qc_var: the trainer's call shape, (B, N) LLRs and targets straight into the QC kernels; qc_var_expand: the same call
    # through index_select expansions to (B, E) (what that shape cost before the per-variable kernels existed)
    for tag, fused, qc in (("qc", True, True), ("qc_var", True, True), ("qc_var_expand", True, True), ("fused", True, False),
                           ("composed", False, False)):
        if tag.startswith("qc") and args.z not in (4, 8, 16, 32):
            continue
        var_shape = tag.startswith("qc_var")
        dec = LDPCNeuralDecoder(code.E, args.iters, 2, output_index_tensor=oidx if var_shape else None, fused=fused, qc=qc).to(dev)
        if tag == "qc_var_expand":
            dec._var_major = False
        xin, yin = (llr, gt_v) if var_shape else (llr_e, gt)

        def infer():
            with torch.no_grad():
                dec(xin, cidx, vidx)

        def train():
            dec.zero_grad(set_to_none=True)
            _, ml = dec(xin, cidx, vidx, yin)
            ml.mean().backward()

        ms_i, ms_t = timed(infer, args.reps), timed(train, args.reps)
        out[tag] = {"infer_ms": ms_i, "infer_cw_per_s": args.batch / ms_i * 1e3,
                    "train_ms": ms_t, "train_cw_per_s": args.batch / ms_t * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
