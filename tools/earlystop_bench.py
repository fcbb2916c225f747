#!/usr/bin/env python
"""Throughput of the reference-default decode (early_stopping=True, max 50 iterations), BG2 Z=32 (or --z 16, the
reference's default lifting factor), 32 768 codewords."""
import sys, time, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder
from ldpc_b200.utils import QCCode
Zarg = int(sys.argv[sys.argv.index("--z") + 1]) if "--z" in sys.argv else 32
code = QCCode.nr_2_0(Zarg)
dev = torch.device("cuda", 0)
B = 32768
llr = torch.empty((B, code.N), dtype=torch.float32, device=dev)
for snr in (-2.0, 0.0):
    _native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, snr, 1, 0, _native.ptr(llr), None))
    for name, dec in (("minsum", MinSumScaledDecoder(code, 50, 0.75, early_stopping=True)),
                      ("minsum check_finite=False", MinSumScaledDecoder(code, 50, 0.75, early_stopping=True, check_finite=False)),
                      ("bp", BeliefPropagationDecoder(code, 50, early_stopping=True)),
                      ("bp path=fast", BeliefPropagationDecoder(code, 50, early_stopping=True, path="fast"))):
        for tag, fn in (("decode(batch-global stop)", lambda: dec.decode(llr)), ("decode_with_iterations(per-codeword)", lambda: dec.decode_with_iterations(llr))):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            extra = r[1] if isinstance(r[1], int) else float(r[1].float().mean())
            print(f"snr {snr} {name} {tag}: {dt*1e3:.1f} ms = {B/dt/1e6:.3f} M cw/s, iterations {extra}")
