"""5G NR rate matching and the punctured / rate-matched LLR layout.

The reference has no rate matching: its loops send the full, un-punctured all-zero codeword of N = cols*Z bits
(trainer.py:86, comparative_evaluation.py:132).  A 5G transmitter never sends that vector: the first 2Z systematic
bits are punctured, filler bits are skipped, E bits are read from a circular buffer at a redundancy-version offset
and interleaved (3GPP TS 38.212 sections 5.3.2, 5.4.2.1, 5.4.2.2).  `RateMatcher` builds those index tables once
(numpy, host) and runs both directions on the device through the C ABI (`ldpc_rate_match`, `ldpc_rate_recover`):

    tx = RateMatcher(code, E, payload_bits=K', rv=0, Qm=2)
    info = tx.pad_info(payload)                    # (B, K') -> (B, K): filler bits are zeros for the encoder
    bits = tx.rate_match(SystematicEncoder(code).encode(info))          # (B, E) transmitted bits
    llr  = tx.rate_recover(channel_llrs)                                  # (B, E) -> (B, N) decoder input
    soft, hard = MinSumScaledDecoder(code, ...).forward(llr)              # hard[:, :K'] is the payload

Layout of the recovered LLR vector (what the decoders see): positions 0..2Z-1 (punctured) and every position that was
not transmitted carry LLR 0; filler positions K'..K-1 carry `filler_llr` (known zeros); a position transmitted more than
once (E larger than the buffer) carries the fp32 sum of its copies in circular-buffer order.
"""
import numpy as np
import torch

from .. import _native

# k0 numerators of TS 38.212 Table 5.4.2.1-2, by number of base-graph columns (BG1: 68 -> N = 66 Zc, BG2: 52 -> 50 Zc)
_K0_NUM = {68: (0, 17, 33, 56), 52: (0, 13, 25, 43)}


def rate_match_tables(code, E, payload_bits=None, rv=0, Qm=1, Ncb=None):
    """(sel [E] int32, base_kind [N] int8 {0: ordinary, 1: punctured, 2: filler}) for the given configuration."""
    Z, N, K = code.Z, code.N, code.K
    Kp = K if payload_bits is None else int(payload_bits)
    Nd = N - 2 * Z                                     # length of the 38.212 codeword d = full[2Z:]
    Ncb = Nd if Ncb is None else int(Ncb)
    E, Qm, rv = int(E), int(Qm), int(rv)
    if not 2 * Z < Kp <= K:
        raise ValueError(f"payload_bits must be in ({2 * Z}, {K}]")
    if not 0 < Ncb <= Nd:
        raise ValueError(f"Ncb must be in (0, {Nd}]")
    if E <= 0 or Qm <= 0 or E % Qm:
        raise ValueError("E must be a positive multiple of Qm")
    if rv not in (0, 1, 2, 3):
        raise ValueError("rv must be 0..3")
    if rv and code.cols not in _K0_NUM:
        raise ValueError("redundancy versions other than 0 are defined for the 5G base graphs (52 or 68 columns) only")
    num = _K0_NUM.get(code.cols, (0, 0, 0, 0))[rv]
    k0 = (num * Ncb // (Nd)) * Z                       # floor(num * Ncb / (Nd/Z * Z)) * Z with Nd = 50 Z or 66 Z
    d = np.arange(Ncb)
    is_null = (d >= Kp - 2 * Z) & (d < K - 2 * Z)      # filler bits inside the circular buffer
    usable = d[~is_null]
    if usable.size == 0:
        raise ValueError("circular buffer holds only filler bits")
    # bit selection: walk the buffer from k0, skip NULLs, wrap, until E bits are taken
    order = np.concatenate([d[k0:], d[:k0]])
    order = order[~is_null[order]]
    e = order[np.arange(E) % order.size]
    # bit interleaving: f[i + j*Qm] = e[i*(E/Qm) + j]
    f = e.reshape(Qm, E // Qm).T.reshape(-1)
    sel = (f + 2 * Z).astype(np.int32)
    kind = np.zeros(N, dtype=np.int8)
    kind[:2 * Z] = 1
    kind[Kp:K] = 2
    return sel, kind


class RateMatcher:
    def __init__(self, code, E, payload_bits=None, rv=0, Qm=1, Ncb=None, filler_llr=1e4):
        """filler_llr: LLR given to the known-zero filler positions.  A large finite value by default: the specialised
        min-sum kernel forms v2c as posterior - own message, which +inf would turn into NaN (pass float('inf') with
        path="exact" decoders for the textbook value)."""
        self.code, self.E = code, int(E)
        self.payload_bits = code.K if payload_bits is None else int(payload_bits)
        self.rv, self.Qm = int(rv), int(Qm)
        self.sel, self.kind = rate_match_tables(code, E, payload_bits, rv, Qm, Ncb)
        # copies of a position are combined in the order the circular buffer produced them (index k of e, i.e. BEFORE the
        # bit interleaver: t = i + j*Qm  <->  k = i*(E/Qm) + j), as a receiver that de-interleaves first would
        t = np.arange(self.E)
        k_of_t = (t % self.Qm) * (self.E // self.Qm) + t // self.Qm
        order = np.lexsort((k_of_t, self.sel))
        counts = np.bincount(self.sel, minlength=code.N)
        self.inv_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.inv_idx = order.astype(np.int32)
        self.base = np.where(self.kind == 2, np.float32(filler_llr), np.float32(0.0)).astype(np.float32)
        self._dev = {}

    def _tables(self, dev):
        key = (dev.type, dev.index)
        if key not in self._dev:
            self._dev[key] = tuple(torch.from_numpy(a).to(dev) for a in (self.sel, self.inv_ptr, self.inv_idx, self.base))
        return self._dev[key]

    @staticmethod
    def _device_of(t):
        if t.is_cuda:
            return t.device
        if not torch.cuda.is_available():
            raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())

    def pad_info(self, payload):
        """(B, K') payload -> (B, K) encoder input with zero filler bits."""
        if payload.shape[1] != self.payload_bits:
            raise ValueError(f"payload must have shape (batch, {self.payload_bits})")
        out = torch.zeros((payload.shape[0], self.code.K), dtype=payload.dtype, device=payload.device)
        out[:, :self.payload_bits] = payload
        return out

    def rate_match(self, codeword):
        """(B, N) codeword bits -> (B, E) transmitted bits (float32 0/1 on the input's device)."""
        if codeword.dim() != 2 or codeword.shape[1] != self.code.N:
            raise ValueError(f"codeword must have shape (batch, {self.code.N})")
        dev = self._device_of(codeword)
        cw = (codeword.to(dev) != 0).to(torch.uint8).contiguous()
        B = cw.shape[0]
        out = torch.empty((B, self.E), dtype=torch.uint8, device=dev)
        if B:
            sel = self._tables(dev)[0]
            with torch.cuda.device(dev):
                _native.check(_native.lib().ldpc_rate_match(_native.ptr(cw), _native.ptr(sel), B, self.code.N, self.E,
                                                            _native.ptr(out), _native.stream_ptr(dev)))
        return out.to(torch.float32).to(codeword.device)

    def rate_recover(self, rx_llr):
        """(B, E) received LLRs (transmission order) -> (B, N) decoder LLRs in the layout described above."""
        if rx_llr.dim() != 2 or rx_llr.shape[1] != self.E:
            raise ValueError(f"rx_llr must have shape (batch, {self.E})")
        dev = self._device_of(rx_llr)
        rx = rx_llr.detach().to(device=dev, dtype=torch.float32).contiguous()
        B = rx.shape[0]
        out = torch.empty((B, self.code.N), dtype=torch.float32, device=dev)
        if B:
            _, ptr, idx, base = self._tables(dev)
            with torch.cuda.device(dev):
                _native.check(_native.lib().ldpc_rate_recover(_native.ptr(rx), _native.ptr(ptr), _native.ptr(idx), _native.ptr(base),
                                                              B, self.code.N, self.E, _native.ptr(out), _native.stream_ptr(dev)))
        return out.to(rx_llr.device)
