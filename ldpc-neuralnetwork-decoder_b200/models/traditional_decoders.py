"""Classic flooding decoders, drop-in for the reference's models/traditional_decoders.py.

  MinSumScaledDecoder       reference :137-284  (decode :177-260)
  BeliefPropagationDecoder  reference :4-134    (decode :42-109)

Same constructor arguments and the same `decode(llr) -> (decoded_bits float32 (B,N),
num_iterations int)` contract, including the reference's batch-global early-stopping rule
(stop at the first iteration after which EVERY codeword of the batch satisfies all checks,
:102-106 / :255-258).  Added, as BASELINE.json's north star asks: `forward(llr)` /
`__call__` returning `(soft LLRs, hard bits)`, construction from `(base_graph, Z)` or a
`QCCode` so no dense H is needed, and `decode_with_iterations` (per-codeword early exit).

`path`: "exact" = the reference's operation order (min-sum beliefs bit-identical to the reference);
"fast" = the kernels specialised for the shipped 5G tables; "auto" (default) = fast where a specialised
kernel exists, else exact.  "auto"/"fast" is NOT reference operation order: min-sum hard decisions
equal the reference-order kernel's on every fixture and on 2^20 frames at the bench point (bench.py
`parity`), soft outputs agree to rounding (1e-4 relative on converged frames); BP hard decisions are
equal on the same 2^20 frames, the inf/NaN pattern differs in ~1e-8 of the beliefs and finite beliefs
near saturation (|L| > 12) by up to percent level (tanh-domain products are ill-conditioned there).
Batches containing non-finite LLRs are routed to the exact kernel automatically.

All arithmetic happens in the CUDA engine (csrc/decode_exact.cuh, csrc/decode_fast.cuh)
behind the C ABI; these classes only marshal tensors.  CPU tensors are staged through the
current CUDA device and results are returned on the input's device; without a GPU every call
raises -- there is no CPU implementation in this package.
"""
import torch

from .. import _native
from ..utils.ldpc_utils import as_code


class _FloodingDecoder:
    _ALGO = None

    def __init__(self, H=None, max_iterations=50, early_stopping=True, base_graph=None, Z=None, path="auto", check_finite=True):
        self.code = as_code(H, base_graph, Z, allow_split=True)     # Z > 32: held as an equivalent code with Z <= 32, renumbered
        self._perm = {}                                              # device -> (engine <- natural, natural <- engine) index tensors
        # check_finite: look for +-inf / NaN LLRs before every decode under path="auto" (one reduction over the batch and a
        # host sync, ~0.1 ms) and send such batches to the reference-order kernel; False skips the look-up for callers
        # that know their LLRs are finite (a demapper's clipped output)
        self.check_finite = bool(check_finite)
        self.H = H
        self.max_iterations = int(max_iterations)
        self.early_stopping = bool(early_stopping)
        if path not in _native.PATHS:
            raise ValueError(f"path must be one of {sorted(_native.PATHS)}")
        self.path = path
        if self.max_iterations < 1:
            raise ValueError("max_iterations must be >= 1")

    # ---- engine call -------------------------------------------------------------------
    def _alpha(self):
        return 1.0

    def _prepare(self, llr):
        if llr.dim() != 2 or llr.shape[1] != self.code.N:
            raise ValueError(f"llr must have shape (batch, {self.code.N}), got {tuple(llr.shape)}")
        if llr.is_cuda:
            dev = llr.device
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        llr_d = llr.detach().to(device=dev, dtype=torch.float32).contiguous()
        if self.code.var_old_of_new is not None:
            llr_d = llr_d.index_select(1, self._perms(dev)[0])
        # path "auto" sends min-sum to the specialised kernel, whose variable update is posterior - own message: with
        # +-inf channel LLRs (hard-decision inputs) that is inf - inf = NaN where the reference's sum over the OTHER
        # checks keeps inf (traditional_decoders.py:235-244).  Such batches take the reference-order kernel.
        self._route = self.path
        if self.path == "auto" and self.check_finite and llr_d.numel():
            if llr_d.data_ptr() % 16 == 0:              # one read-only pass + a 4-byte read-back
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                with torch.cuda.device(dev):
                    _native.check(_native.lib().ldpc_nonfinite_flag(_native.ptr(llr_d), llr_d.numel(), _native.ptr(flag),
                                                                    _native.stream_ptr(dev)))
                bad = bool(int(flag.item()))
            else:                                       # odd view offsets (rows that are not a multiple of 16 bytes)
                bad = not bool(torch.isfinite(llr_d).all())
            if bad:
                self._route = "exact"
        return llr_d, dev

    def _perms(self, dev):
        """Index tensors of a renumbered (Z > 32) code on `dev`: engine order <- natural order, and back."""
        key = (dev.type, dev.index)
        if key not in self._perm:
            self._perm[key] = (torch.from_numpy(self.code.var_old_of_new).to(dev), torch.from_numpy(self.code.var_new_of_old).to(dev))
        return self._perm[key]

    def _natural(self, t, dev):
        """(B, N) engine-order output -> the caller's variable order."""
        if t is None or self.code.var_old_of_new is None or t.dim() != 2 or t.shape[1] != self.code.N:
            return t
        return t.index_select(1, self._perms(dev)[1])

    def _launch(self, llr_d, dev, iters, stop_mode=_native.STOP_FIXED, soft=True, hard_dtype=_native.HARD_F32,
                syndrome=False, iters_out=False, mask=False, path=None):
        B, N = llr_d.shape
        soft_t = torch.empty((B, N), dtype=torch.float32, device=dev) if soft else None
        if hard_dtype == _native.HARD_F32:
            hard_t = torch.empty((B, N), dtype=torch.float32, device=dev)
        elif hard_dtype == _native.HARD_U8:
            hard_t = torch.empty((B, N), dtype=torch.uint8, device=dev)
        else:
            hard_t = torch.empty((B, (N + 31) // 32), dtype=torch.int32, device=dev)
        syn_t = torch.empty(B, dtype=torch.uint8, device=dev) if syndrome else None
        it_t = torch.empty(B, dtype=torch.int32, device=dev) if iters_out else None
        words = (iters + 63) // 64
        mask_t = torch.empty((B, words), dtype=torch.int64, device=dev) if mask else None
        if B == 0:
            return soft_t, hard_t, syn_t, it_t, mask_t
        L = _native.lib()
        h = self.code.handle(dev)
        p = _native.PATHS[path or getattr(self, '_route', self.path)]
        with torch.cuda.device(dev):
            st = _native.stream_ptr(dev)
            if self._ALGO == _native.ALGO_MINSUM:
                rc = L.ldpc_minsum_decode(h, _native.ptr(llr_d), B, iters, self._alpha(), stop_mode, p,
                                          _native.ptr(soft_t), _native.ptr(hard_t), hard_dtype, _native.ptr(syn_t),
                                          _native.ptr(it_t), _native.ptr(mask_t), words, st)
            else:
                rc = L.ldpc_bp_decode(h, _native.ptr(llr_d), B, iters, stop_mode, p, _native.ptr(soft_t),
                                      _native.ptr(hard_t), hard_dtype, _native.ptr(syn_t), _native.ptr(it_t),
                                      _native.ptr(mask_t), words, st)
        _native.check(rc)
        if hard_dtype != _native.HARD_PACKED:                   # packed words are only used internally (validity probes)
            hard_t = self._natural(hard_t, dev)
        return self._natural(soft_t, dev), hard_t, syn_t, it_t, mask_t

    @staticmethod
    def _first_all_valid(mask_t, iters):
        """Index of the first iteration after which every codeword is valid, or None."""
        if mask_t.shape[0] == 0:
            return None
        shifts = torch.arange(64, device=mask_t.device, dtype=torch.int64)
        ok = (((mask_t.unsqueeze(-1) >> shifts) & 1) != 0).all(dim=0).reshape(-1)[:iters]
        idx = torch.nonzero(ok)
        return int(idx[0]) if idx.numel() else None

    def _fast_early_allowed(self):
        """The specialised early-exit kernel may replace the exact validity-mask pass unless path is "exact"."""
        return getattr(self, "_route", self.path) != "exact"

    def _decode_full(self, llr, soft=True):
        """Batch-global early stopping of the reference (:102-106 / :255-258): stop after the first iteration T at
        which EVERY codeword of the batch is valid; return the state after T iterations (T = max_iterations if none).

        Fast route (Z = 32 tables): (A) per-codeword early exit on the specialised kernel gives each codeword's first
        valid iteration; T0 = their maximum is a lower bound of T, and T = max_iterations if a codeword never
        converges.  (B) T0 iterations for the whole batch; if every codeword is valid then, T = T0 exactly.  Only if
        a codeword has left the code again in between (not observed) the exact validity-mask pass decides."""
        llr_d, dev = self._prepare(llr)
        iters = self.max_iterations
        if self.early_stopping and llr_d.shape[0] > 0:
            # two exact passes (docstring), on the specialised kernels where they exist and the policy allows them,
            # otherwise on the exact kernel -- either way far fewer iterations than a max_iterations-long mask pass
            for route in (["fast"] if self._fast_early_allowed() else []) + [self._route]:
                try:
                    _, _, syn, its, _ = self._launch(llr_d, dev, iters, stop_mode=_native.STOP_PER_CODEWORD, soft=False,
                                                     hard_dtype=_native.HARD_PACKED, syndrome=True, iters_out=True, path=route)
                except _native.LdpcError as e:
                    if e.code != _native.ERR_UNSUPPORTED:
                        raise
                    continue
                t0 = int(torch.where(syn.bool().all(), its.max(), torch.tensor(iters, dtype=its.dtype, device=dev)))
                soft_t, hard, syn2, _, _ = self._launch(llr_d, dev, t0, soft=soft, syndrome=True, path=route)
                if t0 == iters or bool(syn2.all()):
                    return (soft_t.to(llr.device) if soft else None), hard.to(llr.device), t0
                break
            soft_t, hard, _, _, mask = self._launch(llr_d, dev, iters, mask=True)
            t = self._first_all_valid(mask, iters)
            if t is not None and t + 1 < iters:
                iters = t + 1
                soft_t, hard, _, _, _ = self._launch(llr_d, dev, iters)
            elif t is not None:
                iters = t + 1
        else:
            soft_t, hard, _, _, _ = self._launch(llr_d, dev, iters, soft=soft)
        return (soft_t.to(llr.device) if soft_t is not None else None), hard.to(llr.device), iters

    # ---- reference API -------------------------------------------------------------------
    def decode(self, llr):
        """(decoded_bits float32 (B,N), num_iterations int), as the reference."""
        _, hard, iters = self._decode_full(llr, soft=False)
        return hard, iters

    # ---- added API -----------------------------------------------------------------------
    def forward(self, llr):
        """(soft posterior LLRs (B,N) fp32, hard bits (B,N) fp32); bit = 1 <=> LLR < 0."""
        soft, hard, _ = self._decode_full(llr)
        return soft, hard

    __call__ = forward

    def decode_with_iterations(self, llr):
        """Per-codeword early exit: (decoded_bits, iterations int32 (B,), syndrome_ok bool (B,)).
        The API `run_comparison_all.py:335-339` expects and the reference never implemented."""
        llr_d, dev = self._prepare(llr)
        stop = _native.STOP_PER_CODEWORD if self.early_stopping else _native.STOP_FIXED
        _, hard, syn, its, _ = self._launch(llr_d, dev, self.max_iterations, stop_mode=stop, soft=False,
                                            syndrome=True, iters_out=True)
        return hard.to(llr.device), its.to(llr.device), syn.to(llr.device).bool()

    def _check_valid_codeword(self, decoded_bits):
        """Bool (B,): all parity checks satisfied (reference :111-134 / :262-284)."""
        bits = decoded_bits.detach()
        if bits.dim() != 2 or bits.shape[1] != self.code.N:
            raise ValueError(f"decoded_bits must have shape (batch, {self.code.N})")
        dev = bits.device if bits.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if bits.dtype == torch.uint8:
            hard, dtype = bits.to(dev).contiguous(), _native.HARD_U8
        else:
            hard, dtype = bits.to(device=dev, dtype=torch.float32).contiguous(), _native.HARD_F32
        if self.code.var_old_of_new is not None:
            hard = hard.index_select(1, self._perms(dev)[0])
        ok = torch.empty(bits.shape[0], dtype=torch.uint8, device=dev)
        if bits.shape[0]:
            with torch.cuda.device(dev):
                _native.check(_native.lib().ldpc_syndrome_check(self.code.handle(dev), _native.ptr(hard), dtype,
                                                                bits.shape[0], _native.ptr(ok), _native.stream_ptr(dev)))
        return ok.bool().to(decoded_bits.device)


class MinSumScaledDecoder(_FloodingDecoder):
    """Scaled min-sum (reference traditional_decoders.py:137-284)."""
    _ALGO = _native.ALGO_MINSUM

    def __init__(self, H=None, max_iterations=50, scaling_factor=0.75, early_stopping=True, base_graph=None, Z=None,
                 path="auto", check_finite=True):
        super().__init__(H, max_iterations, early_stopping, base_graph, Z, path, check_finite)
        self.scaling_factor = scaling_factor

    def _alpha(self):
        return float(self.scaling_factor)


class BeliefPropagationDecoder(_FloodingDecoder):
    """Sum-product (tanh/atanh), unclipped fp32 (reference traditional_decoders.py:4-134)."""
    _ALGO = _native.ALGO_BP
