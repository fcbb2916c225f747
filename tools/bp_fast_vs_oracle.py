import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200 import _native
if os.environ.get("LDPC_LIB"): _native.LIB_PATH = os.environ["LDPC_LIB"]
from ldpc_b200.models import BeliefPropagationDecoder
from ldpc_b200.utils import QCCode
from oracle import oracle
code = QCCode.nr_2_0(32)
for snr, seed in ((-2.0, 3), (4.0, 5), (-3.5, 6)):
    llr = oracle.awgn_llr(None, 256, code.N, snr, seed=seed)
    o = oracle.decode(code.shifts, 32, llr, 10, "bp", threads=os.cpu_count())
    dec = BeliefPropagationDecoder(code, 10, early_stopping=False, path="fast")
    soft, hard = dec.forward(torch.from_numpy(llr).cuda())
    soft, hard = soft.cpu().numpy(), hard.cpu().numpy().astype(np.uint8)
    cls = lambda x: np.where(np.isnan(x), 3, np.where(np.isposinf(x), 1, np.where(np.isneginf(x), 2, 0)))
    fin = np.isfinite(soft) & np.isfinite(o["beliefs"])
    rel = np.abs(soft[fin] - o["beliefs"][fin]) / np.maximum(np.abs(o["beliefs"][fin]), 1.0)
    conv = (o["hard"].sum(axis=1) == 0)
    rmax, rq = (float(rel.max()), float(np.quantile(rel, 0.999))) if rel.size else (0.0, 0.0)
    print(f"snr {snr}: hard mismatches {int((hard != o['hard']).sum())}, class mismatches {int((cls(soft) != cls(o['beliefs'])).sum())}, "
          f"finite {int(fin.sum())} rel max {rmax:.2e} q99.9 {rq:.2e}, frames converged {int(conv.sum())}")
