#!/usr/bin/env python
"""Kernel micro-benchmark for tuning: times ldpc_minsum_decode of a given libldpc_b200.so
variant (built with different -D macros) on LLRs resident in HBM.  Not part of the product."""
import argparse, ctypes as C, os, sys, json
import numpy as np, torch

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--batch", type=int, default=1 << 19)
ap.add_argument("--iters", type=int, nargs="+", default=[10])
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--algo", default="minsum")
a = ap.parse_args()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.utils import QCCode
code = QCCode.nr_2_0(32)
dev = torch.device("cuda", 0)
B = a.batch
llr = torch.empty((B, code.N), dtype=torch.float32, device=dev)
_native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, -2.0, 1234, 0, _native.ptr(llr), None))
hard = torch.empty((B, 52), dtype=torch.int32, device=dev)
shifts = np.ascontiguousarray(code.shifts.reshape(-1))
ref = None
for path in a.libs:
    L = C.CDLL(path)
    for n, (res, args) in _native.PROTOTYPES.items():
        getattr(L, n).restype = res; getattr(L, n).argtypes = args
    h = C.c_void_p()
    assert L.ldpc_code_create(shifts.ctypes.data_as(C.c_void_p), 42, 52, 32, 0, C.byref(h)) == 0
    for it in a.iters:
        def run():
            if a.algo == "minsum":
                rc = L.ldpc_minsum_decode(h, _native.ptr(llr), B, it, 0.75, 0, 0, None, _native.ptr(hard), 2, None, None, None, 0, None)
            else:
                rc = L.ldpc_bp_decode(h, _native.ptr(llr), B, it, 0, 2, None, _native.ptr(hard), 2, None, None, None, 0, None)
            assert rc == 0, L.ldpc_last_error()
        run(); run(); torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        chk = int(hard.sum().item())
        if it == a.iters[0]:
            ref = chk if ref is None else ref
        print(json.dumps({"lib": os.path.basename(path), "iters": it, "ms": round(ms, 3), "Mcw_s": round(B / ms / 1e3, 2),
                          "checksum": chk}), flush=True)
