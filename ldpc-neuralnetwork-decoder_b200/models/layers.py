"""Edge-space layers, drop-in for the reference's models/layers.py.

  CheckLayer      reference :5-66     unscaled min-sum over listed neighbours
  VariableLayer   reference :69-125   llr + sum of listed check messages
  ResidualLayer   reference :128-168  w_ch*llr + c2v + sum_i w_res[i]*prev[i]
  OutputLayer     reference :171-210  sigmoid(final+llr), per-frame max BCE

Same `nn.Module` interfaces and parameter names (`w_ch`, `w_res`); forward and backward of
the gather-type layers run in the engine's kernels (csrc/layers.cuh) through
torch.autograd.Function.  CUDA tensors only.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _native


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("the LDPC engine layers need CUDA tensors (no CPU fallback)")


# ---- neighbour tables, packed once per tensor -----------------------------------------------
# The reference hands every layer call the [E,K] int64 table (utils/ldpc_utils.py:48-58).  The
# kernels read it far faster as k-major uint16 (csrc/layers.cuh IdxU16), so the first call with
# a given tensor packs it on the device (ldpc_neural_pack_index) and later calls reuse the copy;
# the key holds the tensor's address and version counter, the entry keeps the tensor alive.
_PACK_CACHE = {}
_PACK_LIMIT = 16


class _Packed:
    __slots__ = ("src", "table", "symmetric", "_i64", "_sorted")

    def compacted(self):
        """(table, None, cnt): the lists compacted but the columns left in edge order (identity permutation)."""
        compact, _, cnt = sort_plan(self._i64, sort=False)
        return _pack(compact), None, cnt

    def sorted(self):
        """(table [K,E] uint16, perm [E] uint16, cnt [E] uint8) for the engine's sorted-pack kernels."""
        if self._sorted is None:
            compact, perm, cnt = sort_plan(self._i64)
            self._sorted = (_pack(compact), perm.to(torch.int32).to(torch.int16).contiguous(), cnt)
        return self._sorted


def sort_plan(src, sort=True):
    """Host-side plan of the sorted-pack format (pure tensor algebra, any device).  src: [E,K] int64, -1 padded.
    Returns (rows [E,K] int64, perm [E] int64, cnt [E] uint8): row t lists the neighbours of edge perm[t] with the
    valid entries first in the caller's order and unused slots 0; rows are ordered by descending neighbour count
    (stable, so equal counts keep edge order); cnt[t] = number of valid entries of row t."""
    E, K = src.shape
    if K > 255:
        raise ValueError("neighbour tables wider than 255 slots are not supported")
    valid = src >= 0
    order = torch.argsort((~valid).to(torch.int8), dim=1, stable=True)
    compact = torch.gather(src, 1, order)
    cnt = valid.sum(dim=1)
    perm = torch.argsort(cnt, descending=True, stable=True) if sort else torch.arange(E, device=src.device)
    return compact[perm].clamp_min(0).contiguous(), perm, cnt[perm].to(torch.uint8).contiguous()


def _pack(src):
    E, K = src.shape
    out = torch.empty((K, E), dtype=torch.int16, device=src.device)
    with torch.cuda.device(src.device):
        _native.check(_native.lib().ldpc_neural_pack_index(
            _native.ptr(src), E, K, _native.ptr(out), _native.stream_ptr(src.device)))
    return out


def packed_index(idx):
    """-> _Packed(table [K,E] int16 view of the uint16 words, symmetric flag) or None when the
    table cannot be packed (E >= 65535)."""
    E, K = idx.shape
    if E >= 0xFFFF:
        return None
    key = (idx.data_ptr(), idx._version, E, K, idx.dtype, str(idx.device))
    hit = _PACK_CACHE.get(key)
    if hit is None:
        src = idx.to(torch.int64).contiguous()
        out = _pack(src)
        # "n is listed by e  <=>  e is listed by n" (true for the reference's tables: the OTHER edges of the same
        # node).  Then the transpose of the gather is the same gather and backward needs no atomics.
        e_of = torch.arange(E, device=src.device).unsqueeze(1).expand(E, K)[src >= 0]
        n_of = src[src >= 0]
        hit = _Packed()
        hit.src, hit.table, hit._i64, hit._sorted = idx, out, src, None
        hit.symmetric = bool(torch.equal(torch.sort(e_of * E + n_of).values, torch.sort(n_of * E + e_of).values))
        if len(_PACK_CACHE) >= _PACK_LIMIT:
            _PACK_CACHE.clear()
        _PACK_CACHE[key] = hit
    return hit


def _gather_sum(llr, c2v, idx):
    """llr + sum_k c2v[:, idx[:, k]] on the engine (VariableLayer arithmetic, layers.py:78-125)."""
    B, E = c2v.shape
    out = torch.empty_like(c2v)
    pk = packed_index(idx)
    with torch.cuda.device(c2v.device):
        if pk is not None and E * 16 <= 110 * 1024:
            table, perm, cnt = pk.sorted()
            _native.check(_native.lib().ldpc_variable_layer_fwd_sorted(
                _native.ptr(llr), _native.ptr(c2v), _native.ptr(table), table.shape[0], _native.ptr(cnt),
                _native.ptr(perm), None, None, None, 0, B, E, _native.ptr(out), _native.stream_ptr(c2v.device)))
        else:
            idx = idx.to(torch.int64).contiguous()
            _native.check(_native.lib().ldpc_variable_layer_fwd(
                _native.ptr(llr), _native.ptr(c2v), _native.ptr(idx), B, E, idx.shape[1], _native.ptr(out),
                _native.stream_ptr(c2v.device)))
    return out


def _gather_sum_transposed(idx, g):
    """grad_c2v[b, n] = sum over (e, k) with idx[e, k] == n of g[b, e]."""
    B, E = g.shape
    pk = packed_index(idx)
    if pk is not None and pk.symmetric:
        return _gather_sum(torch.zeros_like(g), g, idx)            # same gather, deterministic, no atomics
    idx = idx.to(torch.int64).contiguous()
    gc = torch.empty_like(g)
    with torch.cuda.device(g.device):
        _native.check(_native.lib().ldpc_variable_layer_bwd(
            _native.ptr(idx), _native.ptr(g), B, E, idx.shape[1], _native.ptr(gc), _native.stream_ptr(g.device)))
    return gc


class _CheckFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        _need_cuda(x, idx)
        x = x.detach().to(torch.float32).contiguous()
        B, E = x.shape
        Em, K = idx.shape
        if Em != E:
            raise ValueError("check_index_tensor must have one row per edge")
        out = torch.empty_like(x)
        am = torch.empty((B, E), dtype=torch.int32, device=x.device)
        pk = packed_index(idx)
        ctx.sorted = pk is not None and E * 16 <= 110 * 1024
        with torch.cuda.device(x.device):
            if ctx.sorted:
                # am = edge selected as the minimum (-1: none); backward needs only (x, out, am)
                table, perm, cnt = pk.sorted()
                _native.check(_native.lib().ldpc_check_layer_fwd_sorted(
                    _native.ptr(x), _native.ptr(table), K, _native.ptr(cnt), _native.ptr(perm), B, E, _native.ptr(out),
                    _native.ptr(am), _native.stream_ptr(x.device)))
                ctx.save_for_backward(x, out, am)
            else:
                idx = idx.to(torch.int64).contiguous()
                _native.check(_native.lib().ldpc_check_layer_fwd(
                    _native.ptr(x), _native.ptr(idx), B, E, K, _native.ptr(out), _native.ptr(am),
                    _native.stream_ptr(x.device)))
                ctx.save_for_backward(x, idx, am)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.to(torch.float32).contiguous()
        if ctx.sorted:
            x, out, am = ctx.saved_tensors
            B, E = x.shape
            gx = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _native.check(_native.lib().ldpc_check_layer_bwd_nstar(
                    _native.ptr(x), _native.ptr(out), _native.ptr(am), _native.ptr(g), B, E, _native.ptr(gx),
                    _native.stream_ptr(x.device)))
            return gx, None
        x, idx, am = ctx.saved_tensors
        B, E = x.shape
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().ldpc_check_layer_bwd(
                _native.ptr(x), _native.ptr(idx), _native.ptr(am), _native.ptr(g), B, E, idx.shape[1],
                _native.ptr(gx), _native.stream_ptr(x.device)))
        return gx, None


class _VariableFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, llr, c2v, idx):
        _need_cuda(llr, c2v, idx)
        llr = llr.detach().to(torch.float32).contiguous()
        c2v = c2v.detach().to(torch.float32).contiguous()
        if idx.shape[0] != c2v.shape[1]:
            raise ValueError("var_index_tensor must have one row per edge")
        out = _gather_sum(llr, c2v, idx)
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.to(torch.float32).contiguous()
        return g, _gather_sum_transposed(idx, g), None


class _ResidualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, llr, c2v, w_ch, w_res, *prev):
        _need_cuda(llr, c2v, w_ch, w_res, *prev)
        ts = [t.detach().to(torch.float32).contiguous() for t in (llr, c2v, w_ch, w_res) + tuple(prev)]
        llr_c, c2v_c, wch_c, wres_c, prev_c = ts[0], ts[1], ts[2], ts[3], ts[4:]
        B, E = llr_c.shape
        out = torch.empty_like(llr_c)
        arr = (C.c_void_p * max(len(prev_c), 1))(*[p.data_ptr() for p in prev_c])
        with torch.cuda.device(llr_c.device):
            _native.check(_native.lib().ldpc_residual_layer_fwd(
                _native.ptr(llr_c), _native.ptr(c2v_c), _native.ptr(wch_c), _native.ptr(wres_c), arr, len(prev_c),
                B, E, _native.ptr(out), _native.stream_ptr(llr_c.device)))
        ctx.save_for_backward(llr_c, wch_c, wres_c, *prev_c)
        return out

    @staticmethod
    def backward(ctx, g):
        llr, w_ch, w_res, *prev = ctx.saved_tensors
        # elementwise products / column sums of the upstream gradient: plain tensor algebra
        g_llr = g * w_ch.unsqueeze(0)
        g_wch = (g * llr).sum(dim=0)
        g_wres = torch.zeros_like(w_res)
        g_prev = []
        for i, p in enumerate(prev):
            g_wres[i] = (g * p).sum()
            g_prev.append(g * w_res[i])
        return (g_llr, g, g_wch, g_wres, *g_prev)


class _OutputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, final_llr, llr, gt):
        _need_cuda(final_llr, llr, gt)
        f = final_llr.detach().to(torch.float32).contiguous()
        l = llr.detach().to(torch.float32).contiguous()
        y = gt.detach().to(torch.float32).contiguous() if gt is not None else None
        B, E = f.shape
        soft = torch.empty_like(f)
        ml = torch.empty(B, dtype=torch.float32, device=f.device) if y is not None else None
        am = torch.empty(B, dtype=torch.int32, device=f.device) if y is not None else None
        with torch.cuda.device(f.device):
            _native.check(_native.lib().ldpc_output_layer_fwd(
                _native.ptr(f), _native.ptr(l), _native.ptr(y), B, E, _native.ptr(soft), _native.ptr(ml),
                _native.ptr(am), _native.stream_ptr(f.device)))
        ctx.has_gt = y is not None
        if y is not None:
            ctx.save_for_backward(soft, y, am)
            ctx.mark_non_differentiable(am)
            return soft, ml
        ctx.save_for_backward(soft)
        return soft, None

    @staticmethod
    def backward(ctx, g_soft, g_ml):
        if ctx.has_gt:
            soft, y, am = ctx.saved_tensors
        else:
            (soft,) = ctx.saved_tensors
        gz = torch.zeros_like(soft)
        if g_soft is not None:
            gz = gz + g_soft * soft * (1.0 - soft)
        if ctx.has_gt and g_ml is not None:
            # chain of torch's own backward formulas at the arg-max bit of each frame:
            # BCE' = (s - y) / max(s (1 - s), 1e-12), sigmoid' = s (1 - s)  (vanishes when saturated)
            rows = torch.arange(soft.shape[0], device=soft.device)
            cols = am.long()
            s, t = soft[rows, cols], y[rows, cols]
            gz[rows, cols] += g_ml * (s - t) / torch.clamp_min((1.0 - s) * s, 1e-12) * (s * (1.0 - s))
        return gz, gz, None


class CheckLayer(nn.Module):
    def forward(self, input_tensor, check_index_tensor):
        return _CheckFn.apply(input_tensor, check_index_tensor)


class VariableLayer(nn.Module):
    def forward(self, input_llr, check_messages, var_index_tensor):
        return _VariableFn.apply(input_llr, check_messages, var_index_tensor)


class ResidualLayer(nn.Module):
    def __init__(self, num_nodes, depth_L=2):
        super().__init__()
        self.num_nodes = num_nodes
        self.depth_L = depth_L
        self.w_ch = nn.Parameter(torch.ones(num_nodes))
        self.w_res = nn.Parameter(torch.ones(depth_L))

    def forward(self, input_llr, check_messages, prev_var_messages):
        prev = list(prev_var_messages)[:self.depth_L]        # reference :164-166 ignores i >= depth_L
        return _ResidualFn.apply(input_llr, check_messages, self.w_ch, self.w_res, *prev)


class OutputLayer(nn.Module):
    def forward(self, final_llr, input_llr, ground_truth=None):
        soft, max_loss = _OutputFn.apply(final_llr, input_llr, ground_truth)
        return soft, max_loss
