#!/usr/bin/env python
"""One reference-order (path="exact") min-sum or BP decode of BG2 Z=32, for ncu captures of decode_exact_kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ldpc_b200  # noqa: E402,F401
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder  # noqa: E402
from ldpc_b200.utils import QCCode  # noqa: E402
algo = sys.argv[1] if len(sys.argv) > 1 else "minsum"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
code = QCCode.nr_2_0(32)
llr = torch.randn(B, code.N, device="cuda") * 1.2 + 1.0
cls = MinSumScaledDecoder if algo == "minsum" else BeliefPropagationDecoder
dec = cls(code, max_iterations=10, early_stopping=False, path="exact", check_finite=False)
for _ in range(2):
    dec.forward(llr)
torch.cuda.synchronize()
print("ok")
