#!/usr/bin/env python
"""GNN forward accuracy against the fp64-accumulated oracle (oracle/oracle.py:gnn_forward) on 8 codewords, BG2 Z=32,
5 layers, random init (seed 0): max |soft - ref| / max(|ref|, 1).  Tolerance of the path: 1e-4."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
if os.environ.get("LDPC_LIB"):
    from ldpc_b200 import _native as _n
    _n.LIB_PATH = os.environ["LDPC_LIB"]
from ldpc_b200 import _native
from ldpc_b200.models import create_message_gnn_decoder
from ldpc_b200.utils import QCCode
from oracle import oracle
code = QCCode.nr_2_0(32)
torch.manual_seed(0)
gnn, conv = create_message_gnn_decoder(code, 5, 64, base_graph=code.base_matrix(), Z=32)
dev = torch.device("cuda", 0)
llr = torch.empty((8, code.N), dtype=torch.float32, device=dev)
_native.check(_native.lib().ldpc_awgn_llr(None, 8, code.N, -2.0, 7, 0, _native.ptr(llr), None))
out = gnn(llr)
sd = {k: v.detach().cpu().numpy() for k, v in gnn.state_dict().items()}
ref, _ = oracle.gnn_forward(sd, llr.cpu().numpy(), conv.message_var_index.numpy(), conv.message_check_index.numpy(), gnn._expanded_types(), 5)
rel = np.abs(out[0].cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
print("max_rel_err_soft_vs_oracle", float(rel.max()), "mean", float(rel.mean()))
