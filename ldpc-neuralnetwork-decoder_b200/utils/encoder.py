"""Systematic encoder for the QC codes of the 5G shape.

The reference has no encoder: every training / evaluation loop transmits the all-zero codeword
(trainer.py:86, comparative_evaluation.py:132), which hides anything that is not symmetric in the codeword (the GNN
decoder, the NaN -> bit 0 rule of the unclipped BP).  This module lets the engine transmit real codewords.

Structure used (discovered from the shift table, not assumed): with kb = cols - rows information column blocks,
    H = [ A  B  0 ]      g core rows,   B: g x g blocks of "core parity" columns
        [ C  D  I ]      extension rows, each owning one degree-1 column with shift 0
so  p_core = B^-1 (A s)  and  p_ext = C s + D p_core  over GF(2).  B^-1 (gZ x gZ bits, 128 x 128 at BG2 Z=32) is
computed once on the host by Gaussian elimination; the device kernel (csrc/encode.cuh, `ldpc_encode`) does the
circulant row sums with rotations by address and the dense B^-1 product with popcounts, one warp per codeword.
A table that does not have this shape raises ValueError (no fallback).
"""
import numpy as np
import torch

from .. import _native


def _gf2_inverse(Bm):
    """Inverse of a square 0/1 matrix over GF(2) (Gauss-Jordan); ValueError if singular."""
    n = Bm.shape[0]
    a = np.concatenate([Bm.astype(np.uint8) & 1, np.eye(n, dtype=np.uint8)], axis=1)
    for col in range(n):
        piv = np.nonzero(a[col:, col])[0]
        if piv.size == 0:
            raise ValueError("core parity block of the code is singular over GF(2): no systematic encoder of this shape")
        p = col + int(piv[0])
        if p != col:
            a[[col, p]] = a[[p, col]]
        rows = np.nonzero(a[:, col])[0]
        rows = rows[rows != col]
        a[rows] ^= a[col]
    return a[:, n:]


class SystematicEncoder:
    def __init__(self, code):
        self.code = code
        sh, Z = np.asarray(code.shifts), code.Z
        rows, cols = sh.shape
        kb = cols - rows
        if kb <= 0:
            raise ValueError("code has no information columns")
        deg = (sh >= 0).sum(axis=0)
        ext_of_row = np.full(rows, -1, dtype=np.int32)
        for j in range(kb, cols):
            if deg[j] == 1:
                i = int(np.nonzero(sh[:, j] >= 0)[0][0])
                if sh[i, j] == 0 and ext_of_row[i] < 0:
                    ext_of_row[i] = j
        core_rows = np.nonzero(ext_of_row < 0)[0].astype(np.int32)
        ext_cols = set(int(x) for x in ext_of_row if x >= 0)
        core_cols = np.array([j for j in range(kb, cols) if j not in ext_cols], dtype=np.int32)
        g = len(core_rows)
        if g == 0 or len(core_cols) != g:
            raise ValueError(f"code is not of the shape [A B 0; C D I]: {g} core rows, {len(core_cols)} core parity columns")
        allowed = set(range(kb)) | set(int(c) for c in core_cols)
        for i in range(rows):
            for j in np.nonzero(sh[i] >= 0)[0]:
                if int(j) not in allowed and int(j) != int(ext_of_row[i]):
                    raise ValueError(f"row {i} touches the extension column {int(j)} of another row: not of the shape [A B 0; C D I]")
        Bm = np.zeros((g * Z, g * Z), dtype=np.uint8)
        r = np.arange(Z)
        for a, i in enumerate(core_rows):
            for b, j in enumerate(core_cols):
                s = int(sh[i, j])
                if s >= 0:
                    Bm[a * Z + r, b * Z + (r + s) % Z] = 1          # check i*Z+r <-> variable j*Z+((r+s) mod Z)
        binv = _gf2_inverse(Bm)
        self.g, self.kb = g, kb
        self.words = (g * Z + 31) // 32
        packed = np.zeros((g * Z, self.words), dtype=np.uint32)
        for k in range(g * Z):
            packed[:, k >> 5] |= binv[:, k].astype(np.uint32) << np.uint32(k & 31)
        self.binv_packed = packed
        self.plan = np.concatenate([[g, kb, self.words], core_rows, core_cols, ext_of_row]).astype(np.int32)
        self._dev = {}

    def _tables(self, dev):
        key = (dev.type, dev.index)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.plan).to(dev), torch.from_numpy(self.binv_packed.view(np.int32)).to(dev))
        return self._dev[key]

    def encode(self, info_bits):
        """(B, K) 0/1 values (any dtype, any device) -> codewords (B, N) float32 0/1 on the input's device; the first K
        positions are the information bits (systematic)."""
        squeeze = info_bits.dim() == 1
        s2 = info_bits.unsqueeze(0) if squeeze else info_bits
        code = self.code
        if s2.shape[1] != code.K:
            raise ValueError(f"info_bits must have shape (batch, {code.K}), got {tuple(s2.shape)}")
        if s2.is_cuda:
            dev = s2.device
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        info_u8 = (s2.to(dev) != 0).to(torch.uint8).contiguous()
        B = info_u8.shape[0]
        out = torch.empty((B, code.N), dtype=torch.uint8, device=dev)
        if B:
            plan, binv = self._tables(dev)
            with torch.cuda.device(dev):
                _native.check(_native.lib().ldpc_encode(code.handle(dev), _native.ptr(info_u8), B, _native.ptr(plan), plan.numel(),
                                                        _native.ptr(binv), _native.ptr(out), _native.stream_ptr(dev)))
        cw = out.to(torch.float32).to(info_bits.device)
        return cw.squeeze(0) if squeeze else cw
