#!/usr/bin/env python
"""Throughput of the exact sum-product kernel (path="auto" for BP), BG2 Z=32, 10 iterations, 65 536 codewords."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import BeliefPropagationDecoder, MinSumScaledDecoder
from ldpc_b200.utils import QCCode
code = QCCode.nr_2_0(32)
B = 1 << 16
llr = torch.empty((B, code.N), dtype=torch.float32, device="cuda")
_native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, -2.0, 1, 0, _native.ptr(llr), None))
for name, dec in (("bp exact", BeliefPropagationDecoder(code, 10, early_stopping=False)),
                  ("minsum exact", MinSumScaledDecoder(code, 10, 0.75, early_stopping=False, path="exact"))):
    dec.forward(llr); torch.cuda.synchronize()
    t0 = time.perf_counter(); dec.forward(llr); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, round(B / dt / 1e6, 3), "M cw/s")
