#!/usr/bin/env python
"""Throughput of the table-driven (reference-order) decoder: BG2 Z=32 with path="exact", and BG2's support lifted with
Z = 64 ... 384 (random shifts), which the classic decoders hold as renumbered 32-circulants.  10 iterations, no early stop."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ldpc_b200  # noqa: E402,F401
if os.environ.get("LDPC_LIB"):
    from ldpc_b200 import _native as _n
    _n.LIB_PATH = os.environ["LDPC_LIB"]
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder  # noqa: E402
from ldpc_b200.utils import QCCode  # noqa: E402


def tm(dec, llr):
    dec.forward(llr); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t = time.time(); dec.forward(llr); torch.cuda.synchronize(); ts.append(time.time() - t)
    return sorted(ts)[1]


out = {}
code = QCCode.nr_2_0(32)
llr = torch.randn(1 << 17, code.N, device="cuda") * 1.2 + 1.0
for cls in (MinSumScaledDecoder, BeliefPropagationDecoder):
    dt = tm(cls(code, max_iterations=10, early_stopping=False, path="exact"), llr)
    out[f"{cls.__name__}_bg2_z32_exact_cw_per_s"] = llr.shape[0] / dt
del llr
support = code.shifts >= 0
rng = np.random.default_rng(0)
for Z, B in ((48, 1 << 15), (64, 1 << 15), (96, 1 << 14), (128, 1 << 14), (192, 1 << 13), (384, 1 << 12)):
    base = np.where(support, rng.integers(0, Z, size=support.shape), -1)
    dec = MinSumScaledDecoder(base_graph=torch.from_numpy(base.astype(np.float32)), Z=Z, max_iterations=10, early_stopping=False)
    x = torch.randn(B, 52 * Z, device="cuda") * 1.2 + 1.0
    dt = tm(dec, x)
    out[f"minsum_z{Z}"] = {"Zs": dec.code.Z, "cw_per_s": B / dt, "info_mbit_s": B * 10 * Z / dt / 1e6}
print(json.dumps(out))
