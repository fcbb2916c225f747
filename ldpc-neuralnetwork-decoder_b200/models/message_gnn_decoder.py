"""Message-centred GNN decoder, drop-in for the reference's models/message_gnn_decoder.py.

  MessageGNNLayer            reference :15-152   (parameter container; same names/shapes)
  MessageGNNDecoder          reference :155-353  forward / decode
  TannerToMessageGraph       reference :356-536  message list, mappings, message types
  create_message_gnn_decoder reference :539-582

The `nn.Module` parameter layout is the reference's (`input_embedding.*`,
`gnn_layers.{l}.message_type_embeddings`, `.var_to_check_update.{0,2}.*`,
`.check_to_var_update.{0,2}.*`, `.output_projection.*`, `output_layer.*`), so `state_dict`s
interchange.  The forward pass runs in the engine (csrc/gnn.cuh) on the code's Tanner graph:
the dense E x E adjacency matrices the reference multiplies with are per-node means
 and are never built unless a caller asks `TannerToMessageGraph` for them.

Call convention: `forward(input_llr, message_to_var_mapping, message_types=None,
var_to_check_adjacency=None, check_to_var_adjacency=None, ground_truth=None)` as the reference.
`message_to_var_mapping` must be the 1-D long tensor `[v for (v, c) in converter.messages]`
(`converter.message_var_index`) -- the call the reference intends; its own 2-D one-hot
matrix makes the reference read column 0 (a degenerate decoder) and is rejected
here.  The adjacency arguments are accepted for signature compatibility and ignored: the graph
is the one the decoder was created for.  `forward(llr)` alone is the north star's short form.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native
from ..utils.ldpc_utils import QCCode, as_code


class MessageGNNLayer(nn.Module):
    """Parameters of one GNN layer (reference :22-49)."""

    def __init__(self, num_message_types=1, hidden_dim=64):
        super().__init__()
        self.message_type_embeddings = nn.Parameter(torch.randn(num_message_types, hidden_dim))
        self.var_to_check_update = nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.ReLU(),
                                                 nn.Linear(hidden_dim, hidden_dim))
        self.check_to_var_update = nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.ReLU(),
                                                 nn.Linear(hidden_dim, hidden_dim))
        self.output_projection = nn.Linear(hidden_dim, 1)

    def decode_messages(self, message_features):
        return self.output_projection(message_features).squeeze(-1)


class TannerToMessageGraph:
    """Message graph of a parity-check matrix (reference :356-536), built from the QC tables.

    `messages`, `var_to_messages`, `check_to_messages`, `get_message_types` as the reference;
    `message_var_index` / `message_check_index` are the 1-D index tensors the engine uses.
    The dense attributes (`var_to_check_adjacency`, `check_to_var_adjacency`,
    `message_to_var_mapping`) are materialised lazily, only if read."""

    def __init__(self, H=None, base_graph=None, Z=None):
        self.code = as_code(H, base_graph, Z)
        self.H = H
        self.num_checks, self.num_variables = self.code.M, self.code.N
        chk, var = self.code.edges()
        self._chk, self._var = chk, var
        self.messages = list(zip(var.tolist(), chk.tolist()))
        self.var_to_messages = {i: [] for i in range(self.num_variables)}
        self.check_to_messages = {i: [] for i in range(self.num_checks)}
        for m, (v, c) in enumerate(self.messages):
            self.var_to_messages[v].append(m)
            self.check_to_messages[c].append(m)
        self.message_var_index = torch.from_numpy(var.astype(np.int64))
        self.message_check_index = torch.from_numpy(chk.astype(np.int64))
        self._dense = {}

    def _adjacency(self, seg):
        """D^-1/2 (A+I) D^-1/2 with A = 'shares the node' (reference :423-467): 1/d on the block."""
        seg = torch.from_numpy(seg.astype(np.int64))
        same = (seg.unsqueeze(0) == seg.unsqueeze(1)).to(torch.float32)
        deg = same.sum(dim=1)
        dinv = deg.pow(-0.5)
        return dinv.unsqueeze(1) * same * dinv.unsqueeze(0)

    @property
    def var_to_check_adjacency(self):
        if "av" not in self._dense:
            self._dense["av"] = self._adjacency(self._var)
        return self._dense["av"]

    @property
    def check_to_var_adjacency(self):
        if "ac" not in self._dense:
            self._dense["ac"] = self._adjacency(self._chk)
        return self._dense["ac"]

    @property
    def message_to_var_mapping(self):
        if "m2v" not in self._dense:
            m = torch.zeros((len(self.messages), self.num_variables))
            m[torch.arange(len(self.messages)), self.message_var_index] = 1.0
            self._dense["m2v"] = m
        return self._dense["m2v"]

    def base_edge_types(self, base_graph=None, Z=None):
        """Type index per BASE edge (row-major): rank of its shift among the distinct shifts."""
        if base_graph is None or Z is None:
            return np.zeros(self.code.base_edges, dtype=np.int32), 1
        bg = np.rint(torch.as_tensor(base_graph).cpu().numpy()).astype(np.int64)
        uniq = sorted({int(s) for s in bg.reshape(-1) if s >= 0})
        rank = {s: i for i, s in enumerate(uniq)}
        types = [rank[int(bg[i, j])] if bg[i, j] >= 0 else 0
                 for i in range(self.code.rows) for j in range(self.code.cols) if self.code.shifts[i, j] >= 0]
        return np.asarray(types, dtype=np.int32), max(len(uniq), 1)

    def get_message_types(self, base_graph=None, Z=None):
        """Per-message type indices (reference :488-536)."""
        if base_graph is None or Z is None:
            return torch.zeros(len(self.messages), dtype=torch.long)
        bg = np.rint(torch.as_tensor(base_graph).cpu().numpy()).astype(np.int64)
        uniq = sorted({int(s) for s in bg.reshape(-1) if s >= 0})
        rank = {s: i for i, s in enumerate(uniq)}
        sh = bg[self._chk // Z, self._var // Z]
        return torch.tensor([rank[int(s)] if s >= 0 else 0 for s in sh], dtype=torch.long)


class _GnnLossFn(torch.autograd.Function):
    """loss = mean BCE(sigmoid(soft), gt) of the reference (:313-315), with the engine's backward:
    forward runs ldpc_gnn_forward(training=1) + ldpc_gnn_backward and keeps d(loss)/d(params)."""

    @staticmethod
    def forward(ctx, flat, llr, gt, module):
        dev = llr.device
        h = module._handle(dev)
        L = _native.lib()
        B, N = llr.shape
        params = flat.detach().to(device=dev, dtype=torch.float32).contiguous()
        ws_bytes = L.ldpc_gnn_workspace_bytes(h, B, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        prob = torch.empty((B, N), dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        grad = torch.zeros_like(params)
        gt_f = gt.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            st = _native.stream_ptr(dev)
            _native.check(L.ldpc_gnn_forward(h, _native.ptr(params), _native.ptr(llr), B, None, _native.ptr(prob),
                                             _native.ptr(ws), ws_bytes, 1, st))
            _native.check(L.ldpc_gnn_backward(h, _native.ptr(params), _native.ptr(llr), _native.ptr(gt_f), B,
                                              _native.ptr(loss), _native.ptr(grad), _native.ptr(ws), ws_bytes, st))
        ctx.save_for_backward(grad.to(flat.device))
        ctx.mark_non_differentiable(prob)
        return loss, prob

    @staticmethod
    def backward(ctx, g_loss, g_prob):
        (grad,) = ctx.saved_tensors
        return grad * g_loss.to(grad.device), None, None, None


class MessageGNNDecoder(nn.Module):
    def __init__(self, num_messages, num_iterations=5, hidden_dim=64, num_message_types=1, code=None,
                 base_edge_types=None):
        super().__init__()
        if int(hidden_dim) != 64:
            # message_gnn_decoder.py:162,539 accept any width; the tensor-core kernels (128-row tiles, K = 64/128 MMA
            # shapes, TMEM column map) are built for the width every shipped entry point uses
            raise ValueError(f"MessageGNNDecoder: hidden_dim={hidden_dim} is not supported by the B200 engine; "
                             "its kernels are compiled for hidden_dim=64 (the reference's default)")
        self.num_messages = num_messages
        self.num_iterations = num_iterations
        self.hidden_dim = hidden_dim
        self.num_message_types = num_message_types
        self.input_embedding = nn.Linear(1, hidden_dim)
        self.gnn_layers = nn.ModuleList([MessageGNNLayer(num_message_types, hidden_dim) for _ in range(num_iterations)])
        self.output_layer = nn.Linear(hidden_dim, 1)
        self._code = None
        self._base_types = None
        self._handles = {}
        self._mapping_ok = set()
        if code is not None:
            self.attach_code(code, base_edge_types)

    # ---- graph binding -----------------------------------------------------------------
    def attach_code(self, code, base_edge_types=None):
        """Bind the Tanner graph (a QCCode) and the per-base-edge message types."""
        if code.E != self.num_messages:
            raise ValueError(f"code has {code.E} messages, decoder was built for {self.num_messages}")
        self._code = code
        self._base_types = (np.zeros(code.base_edges, dtype=np.int32) if base_edge_types is None
                            else np.ascontiguousarray(base_edge_types, dtype=np.int32))
        self._handles.clear()
        return self

    def _handle(self, dev):
        if self._code is None:
            raise RuntimeError("MessageGNNDecoder has no Tanner graph: build it with create_message_gnn_decoder(...) "
                               "or call attach_code(QCCode, base_edge_types)")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            out = C.c_void_p()
            _native.check(_native.lib().ldpc_gnn_create(
                self._code.handle(dev), self.num_iterations, self.hidden_dim, self.num_message_types,
                self._base_types.ctypes.data_as(C.c_void_p), C.byref(out)))
            h = _GnnHandle(out)
            self._handles[idx] = h
            if _native.lib().ldpc_gnn_param_count(h.ptr) != sum(p.numel() for p in self.parameters()):
                raise RuntimeError("parameter layout mismatch between the module and the engine")
        return h.ptr

    def flat_parameters(self, dev):
        """All parameters in state_dict order as one fp32 buffer (the engine's layout)."""
        return torch.cat([p.detach().reshape(-1).to(device=dev, dtype=torch.float32) for p in self.parameters()])

    def _check_mapping(self, mapping, message_types, dev):
        key = (id(mapping), None if message_types is None else id(message_types))
        if key in self._mapping_ok:
            return
        if mapping is not None:
            if mapping.dim() != 1:
                raise ValueError("message_to_var_mapping must be the 1-D variable index of every message "
                                 "(converter.message_var_index); the reference's 2-D one-hot matrix makes it read "
                                 "column 0 only (SURVEY.md 3c) and is not supported")
            _, var = self._code.edges()
            if mapping.numel() != var.size or not np.array_equal(mapping.detach().cpu().numpy().astype(np.int64), var):
                raise ValueError("message_to_var_mapping does not describe the decoder's Tanner graph")
        if message_types is not None:
            want = np.repeat(self._expanded_types(), 1)
            if not np.array_equal(message_types.detach().cpu().numpy().astype(np.int64), want):
                raise ValueError("message_types differ from the types the decoder was created with")
        self._mapping_ok.add(key)

    def _expanded_types(self):
        code = self._code
        t = []
        k = 0
        per_row = []
        for i in range(code.rows):
            d = int((code.shifts[i] >= 0).sum())
            per_row.append(self._base_types[k:k + d])
            k += d
        for i in range(code.rows):
            t.append(np.tile(np.clip(per_row[i], 0, self.num_message_types - 1), code.Z))
        return np.concatenate(t).astype(np.int64)

    # ---- engine call -------------------------------------------------------------------
    def _run(self, input_llr):
        if input_llr.dim() != 2 or self._code is not None and input_llr.shape[1] != self._code.N:
            raise ValueError(f"input_llr must have shape (batch, {self._code.N if self._code else 'N'})")
        if input_llr.is_cuda:
            dev = input_llr.device
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        h = self._handle(dev)
        llr = input_llr.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, N = llr.shape
        soft = torch.empty((B, N), dtype=torch.float32, device=dev)
        prob = torch.empty((B, N), dtype=torch.float32, device=dev)
        if B == 0:
            return soft, prob
        L = _native.lib()
        params = self.flat_parameters(dev)
        ws_bytes = L.ldpc_gnn_workspace_bytes(h, B, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.check(L.ldpc_gnn_forward(h, _native.ptr(params), _native.ptr(llr), B, _native.ptr(soft),
                                             _native.ptr(prob), _native.ptr(ws), ws_bytes, 0, _native.stream_ptr(dev)))
        return soft, prob

    def forward(self, input_llr, message_to_var_mapping=None, message_types=None, var_to_check_adjacency=None,
                check_to_var_adjacency=None, ground_truth=None):
        """Probabilities (B,N) -- or (probs, loss) when ground_truth is given -- as the reference.
        With only `input_llr` given, returns the north star's (soft LLRs, hard bits)."""
        short = message_to_var_mapping is None and ground_truth is None
        if self._code is not None:
            self._check_mapping(message_to_var_mapping, message_types, None)
        if ground_truth is not None and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step: engine forward (activations kept) + engine backward; loss.backward() then
            # delivers d(loss)/d(parameter) through the differentiable torch.cat below
            if not input_llr.is_cuda:
                raise RuntimeError("training needs CUDA tensors (no CPU fallback)")
            llr = input_llr.detach().to(torch.float32).contiguous()
            flat = torch.cat([p.reshape(-1) for p in self.parameters()])
            loss, prob = _GnnLossFn.apply(flat, llr, ground_truth, self)
            return prob, loss
        soft, prob = self._run(input_llr)
        soft, prob = soft.to(input_llr.device), prob.to(input_llr.device)
        if short:
            return soft, (prob > 0.5).float()
        if ground_truth is not None:
            # mean BCE on the final output (reference :313-315)
            loss = F.binary_cross_entropy(prob, ground_truth.to(prob.device).float())
            return prob, loss
        return prob

    def decode(self, input_llr, message_to_var_mapping=None, message_types=None, var_to_check_adjacency=None,
               check_to_var_adjacency=None):
        """Hard bits, bit = 1 <=> probability > 0.5 (reference :319-353; note the polarity is the
        opposite of the classic decoders')."""
        if self._code is not None:
            self._check_mapping(message_to_var_mapping, message_types, None)
        _, prob = self._run(input_llr)
        return (prob > 0.5).float().to(input_llr.device)


class _GnnHandle:
    def __init__(self, p):
        self.ptr = p

    def __del__(self):
        try:
            _native.lib().ldpc_gnn_destroy(self.ptr)
        except Exception:
            pass


def create_message_gnn_decoder(H=None, num_iterations=5, hidden_dim=64, base_graph=None, Z=None):
    """(decoder, converter) as the reference (:539-582).  `H` may be a dense matrix, a QCCode, or
    omitted when (base_graph, Z) are given."""
    if H is None or isinstance(H, QCCode):
        converter = TannerToMessageGraph(H, base_graph, Z)
    else:
        converter = TannerToMessageGraph(QCCode.from_dense(H, Z))
    base_types, num_types = converter.base_edge_types(base_graph, Z)
    decoder = MessageGNNDecoder(num_messages=len(converter.messages), num_iterations=num_iterations,
                                hidden_dim=hidden_dim, num_message_types=num_types, code=converter.code,
                                base_edge_types=base_types)
    return decoder, converter
