"""Unrolled neural min-sum decoder in edge space.

The reference imports `LDPCNeuralDecoder` from `models/decoder.py` (main.py:6,68-72,
run_comparison.py:18,85-89, training/comparative_evaluation.py:6) but that file is NOT in
the repository; what exists is
  * the four layers it was built from, models/layers.py:5-210 (mirrored in layers.py here),
  * its prototype `LDPCDecoderResidual`, EE4002R_2025.ipynb cell 11 (`forward`,
    `variable_layer_update`),
  * its call sites: `LDPCNeuralDecoder(num_nodes=E, num_iterations=I, depth_L=L)`,
    `decoder(llrs, check_index_tensor, var_index_tensor, ground_truth) -> (soft_bits, loss)`
    with `loss.mean().backward()` (training/trainer.py:102-107,180-187) and
    `decoder.decode(llrs, check_index_tensor, var_index_tensor) -> hard bits`
    (trainer.py:245, comparative_evaluation.py:206).
This class is the composition those three sources imply, written ONLY in terms of the
reference's layer semantics so that it can be pinned against the reference's own layer
classes (fixture tests/golden/neural_decoder_z4.npz):

    x_0 = llr_e                                   (edge copy of the channel LLRs)
    for l in 0 .. I-1:
        c2v = CheckLayer(x_l, check_index_tensor)                       layers.py:14-66
        if l == I-1: break
        s   = VariableLayer(0, c2v, var_index_tensor)                   layers.py:78-125
        x_{l+1} = ResidualLayer(llr_e, s, [x_l, x_{l-1}, ...][:L])      layers.py:143-168
                                            (only iterations >= 1 are in the queue, as in
                                             the notebook: no residual in the first update)
    final = VariableLayer(c2v, c2v, var_index_tensor)   = sum of ALL check messages of the
                                                          edge's variable
    soft, max_loss = OutputLayer(final, llr_e, ground_truth_e)          layers.py:180-210

On the 5G BG2 Z=32 code with the index tensors of create_LLR_mapping (checked once per tensor) both inference and
training run on the QC-structured kernels (csrc/neural_qc_kernel.cuh: no index tables, state in Tensor Memory;
training = one forward kernel that saves the CheckLayer inputs + one backward kernel); `qc=False` disables that.
Otherwise two execution paths, both CUDA, bit-identical to each other and to the QC forward:
  * no gradient needed (`torch.no_grad()`, `.decode()`, frozen parameters): the WHOLE decoder
    is one kernel (`ldpc_neural_decode`, csrc/neural.cuh) that keeps the messages of a
    codeword in shared memory across all iterations; the index tensors are packed once per
    code to k-major uint16 (`ldpc_neural_pack_index`);
  * training: per-layer kernels under torch.autograd.Function wrappers, with the variable +
    residual pair as ONE kernel (`ldpc_neural_variable_layer_fwd`, csrc/layers.cuh).
`fused=False` keeps the literal four-layer composition on both paths (the tests use it to
show the identity).  CUDA tensors only.

Variable-space I/O: the reference's training loop feeds (B, N) LLRs and compares (B, N)
bits, the layers work on (B, E).  Pass `output_index_tensor` (the (1, E) "variable of each
edge" row that `create_LLR_mapping` returns, utils/ldpc_utils.py:93) to the constructor and
the decoder gathers LLRs / ground truth to edges on the way in and returns one soft value
per variable (taken at the variable's first edge) on the way out; without it, inputs and
outputs are edge-space tensors exactly as the layers define them.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _native
from .layers import (CheckLayer, VariableLayer, ResidualLayer, OutputLayer, _need_cuda, packed_index,
                     _gather_sum_transposed)


class _NeuralVariableFn(torch.autograd.Function):
    """out = w_ch*llr + sum_k c2v[idx] + sum_i w_res[i]*prev[i] in one kernel."""

    @staticmethod
    def forward(ctx, llr, c2v, idx, w_ch, w_res, *prev):
        _need_cuda(llr, c2v, idx, w_ch, w_res, *prev)
        ts = [t.detach().to(torch.float32).contiguous() for t in (llr, c2v, w_ch, w_res) + tuple(prev)]
        llr_c, c2v_c, wch_c, wres_c, prev_c = ts[0], ts[1], ts[2], ts[3], ts[4:]
        B, E = llr_c.shape
        if idx.shape[0] != E:
            raise ValueError("var_index_tensor must have one row per edge")
        out = torch.empty_like(llr_c)
        arr = (C.c_void_p * max(len(prev_c), 1))(*[p.data_ptr() for p in prev_c])
        pk = packed_index(idx)
        with torch.cuda.device(llr_c.device):
            if pk is not None and E * 16 <= 110 * 1024:
                table, perm, cnt = pk.sorted()
                _native.check(_native.lib().ldpc_variable_layer_fwd_sorted(
                    _native.ptr(llr_c), _native.ptr(c2v_c), _native.ptr(table), table.shape[0], _native.ptr(cnt),
                    _native.ptr(perm), _native.ptr(wch_c), _native.ptr(wres_c), arr, len(prev_c), B, E,
                    _native.ptr(out), _native.stream_ptr(llr_c.device)))
            else:
                idx = idx.to(torch.int64).contiguous()
                _native.check(_native.lib().ldpc_neural_variable_layer_fwd(
                    _native.ptr(llr_c), _native.ptr(c2v_c), _native.ptr(idx), _native.ptr(wch_c), _native.ptr(wres_c),
                    arr, len(prev_c), B, E, idx.shape[1], _native.ptr(out), _native.stream_ptr(llr_c.device)))
        ctx.save_for_backward(llr_c, idx, wch_c, wres_c, *prev_c)
        return out

    @staticmethod
    def backward(ctx, g):
        llr, idx, w_ch, w_res, *prev = ctx.saved_tensors
        g = g.to(torch.float32).contiguous()
        g_c2v = _gather_sum_transposed(idx, g)
        g_wres = torch.zeros_like(w_res)
        g_prev = []
        for i, p in enumerate(prev):
            g_wres[i] = (g * p).sum()
            g_prev.append(g * w_res[i])
        return (g * w_ch.unsqueeze(0), g_c2v, None, (g * llr).sum(dim=0), g_wres, *g_prev)


# ---- QC-structured path (csrc/neural_qc.cuh) -------------------------------------------------------------------------
# The kernel implies the neighbour tables from the base graph, so it may only run when the caller's tables ARE the ones
# create_LLR_mapping produces for that code.  Checked once per (tensor, version); canonical tables built once per device.
_QC_CANON = {}      # (device index, Z) -> (QCCode, check table, var table, edge -> variable)
_QC_SEEN = []       # [(check tensor, its version, var tensor, its version, ok)], most recent first; the tensors are held so that
                    # their addresses cannot be reused by other data while the verdict is cached


_QC_LIFTS = (32, 16, 8, 4)     # lifting sizes the QC-structured kernels are compiled for (16 = the reference's default --lifting_factor)


def _qc_code_for(check_index_tensor, var_index_tensor, num_nodes):
    from ..utils.ldpc_utils import QCCode, create_LLR_mapping
    Z = num_nodes // 197
    if Z not in _QC_LIFTS or num_nodes != 197 * Z or tuple(check_index_tensor.shape) != (num_nodes, 9) \
            or tuple(var_index_tensor.shape) != (num_nodes, 22):
        return None
    dev = check_index_tensor.device
    canon = _QC_CANON.get((dev.index, Z))
    if canon is None:
        code = QCCode.nr_2_0(Z)
        _, c, v, o = create_LLR_mapping(code.dense().T)
        canon = _QC_CANON[(dev.index, Z)] = (code, c.to(dev), v.to(dev), torch.as_tensor(o).reshape(-1).to(torch.int64).to(dev))
    for i, (ct, cv, vt, vv, ok) in enumerate(_QC_SEEN):
        if ct is check_index_tensor and vt is var_index_tensor and cv == ct._version and vv == vt._version:
            if i:
                _QC_SEEN.insert(0, _QC_SEEN.pop(i))
            return canon[0] if ok else None
    ok = bool(torch.equal(check_index_tensor.to(torch.int64), canon[1]) and torch.equal(var_index_tensor.to(torch.int64), canon[2]))
    _QC_SEEN.insert(0, (check_index_tensor, check_index_tensor._version, var_index_tensor, var_index_tensor._version, ok))
    del _QC_SEEN[8:]
    return canon[0] if ok else None


class _NeuralQcTrainFn(torch.autograd.Function):
    """The whole decoder as ONE forward and ONE backward kernel on the QC structure (csrc/neural_qc_kernel.cuh): the
    call shape of training/trainer.py:95-110 -- `soft, loss = decoder(llr, cidx, vidx, gt); loss.mean().backward()`.
    Differentiable output: max_loss (w.r.t. w_ch and w_res).  `soft` is returned for monitoring and is NOT part of the
    graph on this path (a loss built on `soft`, or gradients w.r.t. the LLRs, take the per-layer path instead)."""

    @staticmethod
    def forward(ctx, llr_e, gt_e, w_ch, w_res, code, iters, depth_L):
        llr_c = llr_e.detach().to(torch.float32).contiguous()
        y = gt_e.detach().to(torch.float32).contiguous()
        wch = w_ch.detach().to(torch.float32).contiguous()
        wres = w_res.detach().to(torch.float32).contiguous()
        B, E = llr_c.shape
        dev = llr_c.device
        soft = torch.empty_like(llr_c)
        ml = torch.empty(B, dtype=torch.float32, device=dev)
        am = torch.empty(B, dtype=torch.int32, device=dev)
        save_x = torch.empty((iters, B, E), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_neural_decode_qc(
                code.handle(dev), _native.ptr(llr_c), _native.ptr(wch), _native.ptr(wres), depth_L, iters, B, _native.ptr(y),
                _native.ptr(soft), _native.ptr(ml), _native.ptr(save_x), _native.ptr(am), _native.stream_ptr(dev)))
        ctx.save_for_backward(save_x, soft, y, am, wres)
        ctx.code, ctx.iters, ctx.depth_L, ctx.E = code, iters, depth_L, E
        ctx.mark_non_differentiable(soft)
        return soft, ml

    @staticmethod
    def backward(ctx, _g_soft, g_ml):
        save_x, soft, y, am, wres = ctx.saved_tensors
        dev = soft.device
        B = soft.shape[0]
        g = (g_ml if g_ml is not None else torch.zeros(B, device=dev)).to(torch.float32).contiguous()
        g_wch = torch.zeros(ctx.E, dtype=torch.float32, device=dev)
        g_wres = torch.zeros(max(ctx.depth_L, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_neural_backward_qc(
                ctx.code.handle(dev), _native.ptr(save_x), _native.ptr(soft), _native.ptr(y), _native.ptr(am), _native.ptr(g),
                _native.ptr(wres), ctx.depth_L, ctx.iters, B, _native.ptr(g_wch), _native.ptr(g_wres), _native.stream_ptr(dev)))
        return None, None, g_wch, g_wres[:ctx.depth_L], None, None, None


class _NeuralQcVarTrainFn(torch.autograd.Function):
    """_NeuralQcTrainFn with per-variable LLRs and targets, the shape the trainer holds them in (trainer.py:95-110): the kernels
    expand a variable's value to its edges on chip (ldpc_neural_decode_qc_var), so no (B, E) tensor is built -- neither the two
    index_select expansions nor the (B, E) soft output.  Bit-identical to the edge-space function on the expanded arrays."""

    @staticmethod
    def forward(ctx, llr_v, gt_v, w_ch, w_res, code, iters, depth_L):
        llr_c = llr_v.detach().to(torch.float32).contiguous()
        y = gt_v.detach().to(torch.float32).contiguous()
        wch = w_ch.detach().to(torch.float32).contiguous()
        wres = w_res.detach().to(torch.float32).contiguous()
        B = llr_c.shape[0]
        E = wch.numel()
        dev = llr_c.device
        soft = torch.empty_like(llr_c)
        ml = torch.empty(B, dtype=torch.float32, device=dev)
        am = torch.empty(B, dtype=torch.int32, device=dev)
        star = torch.empty((B, 2), dtype=torch.float32, device=dev)
        save_x = torch.empty((iters, B, E), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_neural_decode_qc_var(
                code.handle(dev), _native.ptr(llr_c), _native.ptr(wch), _native.ptr(wres), depth_L, iters, B, _native.ptr(y),
                _native.ptr(soft), _native.ptr(ml), _native.ptr(save_x), _native.ptr(am), _native.ptr(star), _native.stream_ptr(dev)))
        ctx.save_for_backward(save_x, star, am, wres)
        ctx.code, ctx.iters, ctx.depth_L, ctx.E = code, iters, depth_L, E
        ctx.mark_non_differentiable(soft)
        return soft, ml

    @staticmethod
    def backward(ctx, _g_soft, g_ml):
        save_x, star, am, wres = ctx.saved_tensors
        dev = star.device
        B = star.shape[0]
        g = (g_ml if g_ml is not None else torch.zeros(B, device=dev)).to(torch.float32).contiguous()
        g_wch = torch.zeros(ctx.E, dtype=torch.float32, device=dev)
        g_wres = torch.zeros(max(ctx.depth_L, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_neural_backward_qc_var(
                ctx.code.handle(dev), _native.ptr(save_x), _native.ptr(star), _native.ptr(am), _native.ptr(g),
                _native.ptr(wres), ctx.depth_L, ctx.iters, B, _native.ptr(g_wch), _native.ptr(g_wres), _native.stream_ptr(dev)))
        return None, None, g_wch, g_wres[:ctx.depth_L], None, None, None


class LDPCNeuralDecoder(nn.Module):
    def __init__(self, num_nodes, num_iterations=5, depth_L=2, output_index_tensor=None, fused=True, qc=True):
        super().__init__()
        if num_iterations < 1:
            raise ValueError("num_iterations must be >= 1")
        self.num_nodes = int(num_nodes)
        self.num_iterations = int(num_iterations)
        self.depth_L = int(depth_L)
        self.fused = bool(fused)
        self.qc = bool(qc)          # allow the QC-structured kernels when the index tensors are the 5G BG2 Z=32 tables
        self.check_layer = CheckLayer()
        self.variable_layer = VariableLayer()
        self.residual_layer = ResidualLayer(self.num_nodes, self.depth_L)   # owns w_ch (E,), w_res (L,)
        self.output_layer = OutputLayer()
        if output_index_tensor is not None:
            m = torch.as_tensor(output_index_tensor).reshape(-1).to(torch.int64)
            if m.numel() != self.num_nodes:
                raise ValueError("output_index_tensor must name the variable of each of the num_nodes edges")
            n_var = int(m.max()) + 1
            first = torch.full((n_var,), self.num_nodes, dtype=torch.int64)
            first.scatter_reduce_(0, m, torch.arange(self.num_nodes), reduce="amin")
            if int(first.max()) >= self.num_nodes:
                raise ValueError("output_index_tensor leaves a variable without an edge")
            self.register_buffer("edge_to_var", m, persistent=False)
            self.register_buffer("var_first_edge", first, persistent=False)
            # create_LLR_mapping's numbering (edges of a variable consecutive, variables ascending): the QC kernels can then
            # take (B, N) LLRs / targets and expand them on chip
            self._var_major = bool(torch.equal(m, torch.sort(m).values))
        else:
            self.edge_to_var = None
            self.var_first_edge = None
            self._var_major = False
        self._etv_canonical = None      # edge_to_var equals the 5G BG2 Z=32 mapping (checked once, on the first per-variable call)

    # -- helpers ---------------------------------------------------------------------------
    def _to_edges(self, t):
        if t is None:
            return None
        if t.dim() != 2:
            raise ValueError("expected a (batch, num_nodes) or (batch, num_variables) tensor")
        if t.shape[1] == self.num_nodes:
            return t
        if self.edge_to_var is not None and t.shape[1] == self.var_first_edge.numel():
            return t.index_select(1, self.edge_to_var.to(t.device))
        raise ValueError(f"tensor has {t.shape[1]} columns; expected {self.num_nodes} edges"
                         + ("" if self.edge_to_var is None else f" or {self.var_first_edge.numel()} variables"))

    def _weights(self):
        """(w_ch (E,), w_res (L,)) the iterations use; TiedNeuralLDPCDecoder expands its per-base-edge weights here."""
        return self.residual_layer.w_ch, self.residual_layer.w_res

    def _messages(self, llr_e, check_index_tensor, var_index_tensor):
        """Last check-to-variable messages after num_iterations unrolled iterations."""
        res = self.residual_layer
        w_ch_t, w_res_t = self._weights()
        queue = []                                            # most recent first
        x = llr_e
        c2v = None
        for l in range(self.num_iterations):
            c2v = self.check_layer(x, check_index_tensor)
            if l == self.num_iterations - 1:
                break
            prev = queue[:self.depth_L]
            if self.fused:
                x = _NeuralVariableFn.apply(llr_e, c2v, var_index_tensor, w_ch_t, w_res_t[:len(prev)], *prev)
            else:
                s = self.variable_layer(torch.zeros_like(c2v), c2v, var_index_tensor)
                if w_ch_t is res.w_ch:
                    x = res(llr_e, s, prev)
                else:                      # ResidualLayer.forward (layers.py:143-168) with the expanded weights
                    x = llr_e * w_ch_t.unsqueeze(0) + s
                    for i, pm in enumerate(prev):
                        x = x + w_res_t[i] * pm
            queue.insert(0, x)
        return c2v

    def _forward_one_kernel(self, llr_e, check_index_tensor, var_index_tensor, gt_e):
        res = self.residual_layer
        llr_c = llr_e.detach().to(torch.float32).contiguous()
        B, E = llr_c.shape
        if check_index_tensor.shape[0] != E or var_index_tensor.shape[0] != E:
            raise ValueError("index tensors must have one row per edge")
        code = _qc_code_for(check_index_tensor, var_index_tensor, E) if (self.qc and self.depth_L <= 2) else None
        if code is not None:
            # the tables are create_LLR_mapping's for 5G BG2 Z=32: neighbours come from the base graph (no index loads)
            w_ch_t, w_res_t = self._weights()
            w_ch = w_ch_t.detach().to(torch.float32).contiguous()
            w_res = w_res_t.detach().to(torch.float32).contiguous()
            y = gt_e.detach().to(torch.float32).contiguous() if gt_e is not None else None
            soft = torch.empty_like(llr_c)
            ml = torch.empty(B, dtype=torch.float32, device=llr_c.device) if y is not None else None
            with torch.cuda.device(llr_c.device):
                _native.check(_native.lib().ldpc_neural_decode_qc(
                    code.handle(llr_c.device), _native.ptr(llr_c), _native.ptr(w_ch), _native.ptr(w_res), self.depth_L,
                    self.num_iterations, B, _native.ptr(y), _native.ptr(soft), _native.ptr(ml), None, None,
                    _native.stream_ptr(llr_c.device)))
            return soft, ml
        cp, cperm, ccnt = packed_index(check_index_tensor).sorted()
        vp, vperm, vcnt = packed_index(var_index_tensor).sorted()
        w_ch_t, w_res_t = self._weights()
        w_ch = w_ch_t.detach().to(torch.float32).contiguous()
        w_res = w_res_t.detach().to(torch.float32).contiguous()
        y = gt_e.detach().to(torch.float32).contiguous() if gt_e is not None else None
        soft = torch.empty_like(llr_c)
        ml = torch.empty(B, dtype=torch.float32, device=llr_c.device) if y is not None else None
        with torch.cuda.device(llr_c.device):
            _native.check(_native.lib().ldpc_neural_decode(
                _native.ptr(llr_c), _native.ptr(cp), cp.shape[0], _native.ptr(ccnt), _native.ptr(cperm),
                _native.ptr(vp), vp.shape[0], _native.ptr(vcnt), _native.ptr(vperm), _native.ptr(w_ch),
                _native.ptr(w_res), self.depth_L, self.num_iterations, B, E, _native.ptr(y), _native.ptr(soft),
                _native.ptr(ml), _native.stream_ptr(llr_c.device)))
        return soft, ml

    def _forward_per_variable(self, llr_v, check_index_tensor, var_index_tensor, gt_v):
        """(B, N) LLRs (and targets) straight into the QC-structured kernels; None when this call cannot take that path (tables
        that are not the 5G BG2 Z=32 ones, targets in edge space, gradients w.r.t. the LLRs, no targets while training)."""
        n_var = self.var_first_edge.numel()
        if llr_v.dim() != 2 or llr_v.shape[1] != n_var or n_var * 197 != 52 * self.num_nodes:
            return None
        if gt_v is not None and tuple(gt_v.shape) != tuple(llr_v.shape):
            return None
        if check_index_tensor.shape[0] != self.num_nodes or var_index_tensor.shape[0] != self.num_nodes:
            return None
        needs_grad = torch.is_grad_enabled() and (llr_v.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad and (gt_v is None or llr_v.requires_grad):
            return None
        code = _qc_code_for(check_index_tensor, var_index_tensor, self.num_nodes)
        if code is None:
            return None
        if self._etv_canonical is None:
            dev = check_index_tensor.device
            self._etv_canonical = bool(torch.equal(self.edge_to_var.to(dev), _QC_CANON[(dev.index, code.Z)][3]))
        if not self._etv_canonical:
            return None
        w_ch_t, w_res_t = self._weights()
        if needs_grad:
            return _NeuralQcVarTrainFn.apply(llr_v, gt_v, w_ch_t, w_res_t, code, self.num_iterations, self.depth_L)
        llr_c = llr_v.detach().to(torch.float32).contiguous()
        w_ch = w_ch_t.detach().to(torch.float32).contiguous()
        w_res = w_res_t.detach().to(torch.float32).contiguous()
        y = gt_v.detach().to(torch.float32).contiguous() if gt_v is not None else None
        B = llr_c.shape[0]
        soft = torch.empty_like(llr_c)
        ml = torch.empty(B, dtype=torch.float32, device=llr_c.device) if y is not None else None
        with torch.cuda.device(llr_c.device):
            _native.check(_native.lib().ldpc_neural_decode_qc_var(
                code.handle(llr_c.device), _native.ptr(llr_c), _native.ptr(w_ch), _native.ptr(w_res), self.depth_L,
                self.num_iterations, B, _native.ptr(y), _native.ptr(soft), _native.ptr(ml), None, None, None,
                _native.stream_ptr(llr_c.device)))
        return soft, ml

    # -- reference-shaped API ----------------------------------------------------------------
    def forward(self, input_llr, check_index_tensor, var_index_tensor, ground_truth=None):
        """-> (soft_bits, max_loss | None); trainer.py:102,180 call shape."""
        _need_cuda(input_llr, check_index_tensor, var_index_tensor, ground_truth)
        per_variable = self.edge_to_var is not None and input_llr.shape[1] != self.num_nodes
        if per_variable and self._var_major and self.fused and self.qc and self.depth_L <= 2:
            out = self._forward_per_variable(input_llr, check_index_tensor, var_index_tensor, ground_truth)
            if out is not None:
                return out
        llr_e = self._to_edges(input_llr).to(torch.float32)
        gt_e = self._to_edges(ground_truth)
        needs_grad = torch.is_grad_enabled() and (llr_e.requires_grad or any(
            p.requires_grad for p in self.parameters()))
        code = None
        if (self.fused and self.qc and needs_grad and gt_e is not None and self.depth_L <= 2 and not llr_e.requires_grad
                and check_index_tensor.shape[0] == self.num_nodes and var_index_tensor.shape[0] == self.num_nodes):
            code = _qc_code_for(check_index_tensor, var_index_tensor, self.num_nodes)
        if code is not None:
            # training on the QC structure: one forward kernel that saves the CheckLayer inputs + one backward kernel
            w_ch_t, w_res_t = self._weights()
            soft, max_loss = _NeuralQcTrainFn.apply(llr_e, gt_e, w_ch_t, w_res_t, code, self.num_iterations, self.depth_L)
        elif self.fused and not needs_grad and self.num_nodes < 0xFFFF and self.depth_L <= 4:
            soft, max_loss = self._forward_one_kernel(llr_e, check_index_tensor, var_index_tensor, gt_e)
        else:
            c2v = self._messages(llr_e, check_index_tensor, var_index_tensor)
            final = self.variable_layer(c2v, c2v, var_index_tensor)  # sum over all checks of the variable
            soft, max_loss = self.output_layer(final, llr_e, gt_e)
        if per_variable:
            soft = soft.index_select(1, self.var_first_edge.to(soft.device))
        return soft, max_loss

    @torch.no_grad()
    def decode(self, input_llr, check_index_tensor, var_index_tensor):
        """Hard decisions with the reference's rule `soft_bits > 0.5` (trainer.py:186)."""
        soft, _ = self.forward(input_llr, check_index_tensor, var_index_tensor)
        return (soft > 0.5).float()


class TiedNeuralLDPCDecoder(LDPCNeuralDecoder):
    """`TiedNeuralLDPCDecoder(base_graph, Z, num_iterations, depth_L)` -- the second class main.py:6,73-79 and
    run_comparison_all.py:21 import from the reference's missing models/decoder.py.  No definition, prototype or test of it
    exists in the reference; its constructor (a base graph and a lifting factor instead of an edge count) and the paper the
    project follows (Deep Neural Network Based Decoding of Short 5G LDPC Codes: "the weights are tied over the Z copies of
    a base-graph edge") fix what it must be: LDPCNeuralDecoder whose channel weights are ONE trainable value per base-graph
    edge, shared by its Z lifted edges.  Everything else -- layers, composition, call shape -- is the parent's, so it
    runs on the same kernels (the QC-structured ones at BG2 Z=32); the gradient of a tied weight is the sum over its Z
    edges (autograd of the expansion).

    The index tensors may be omitted in forward()/decode(): the decoder owns the code and builds create_LLR_mapping's
    tables once (a reference-style caller passing its own tables gets the usual checks)."""

    def __init__(self, base_graph, Z, num_iterations=5, depth_L=2, fused=True, qc=True):
        from ..utils.ldpc_utils import QCCode, create_LLR_mapping
        code = base_graph if isinstance(base_graph, QCCode) else QCCode.from_base_matrix(base_graph, Z)
        _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
        super().__init__(code.E, num_iterations, depth_L, output_index_tensor=oidx, fused=fused, qc=qc)
        self.code = code
        # base edge (cell) of every lifted edge in the variable-major numbering: edges of variable j*Z + r are consecutive,
        # in ascending base row, so the k-th edge of a variable of base column j is cell D_j + k
        deg = (code.shifts >= 0).sum(axis=0)
        d0 = [0]
        for d in deg[:-1]:
            d0.append(d0[-1] + int(d))
        cell = torch.cat([torch.arange(int(deg[j])).repeat(code.Z) + d0[j] for j in range(code.cols)]).to(torch.int64)
        self.register_buffer("edge_cell", cell, persistent=False)
        self.register_buffer("check_index_tensor", cidx, persistent=False)
        self.register_buffer("var_index_tensor", vidx, persistent=False)
        # the parent's per-edge weight is replaced by the tied one (the parameter list holds the tied weights only)
        del self.residual_layer.w_ch
        self.residual_layer.w_ch = None
        self.w_ch_tied = nn.Parameter(torch.ones(code.base_edges))

    def _weights(self):
        return self.w_ch_tied.index_select(0, self.edge_cell), self.residual_layer.w_res

    def forward(self, input_llr, check_index_tensor=None, var_index_tensor=None, ground_truth=None):
        c = self.check_index_tensor if check_index_tensor is None else check_index_tensor
        v = self.var_index_tensor if var_index_tensor is None else var_index_tensor
        return super().forward(input_llr, c, v, ground_truth)

    @torch.no_grad()
    def decode(self, input_llr, check_index_tensor=None, var_index_tensor=None):
        soft, _ = self.forward(input_llr, check_index_tensor, var_index_tensor)
        return (soft > 0.5).float()
