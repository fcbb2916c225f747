"""Code construction: base-graph shift tables instead of a dense H.

Mirror of the reference's ldpc_neural_decoder/utils/ldpc_utils.py (same function names,
arguments and return values) plus `QCCode`, the object the engine actually works with:

  load_base_matrix      ldpc_utils.py:127-146   text file -> (rows, cols) float tensor, -1 = empty
  expand_base_matrix    ldpc_utils.py:97-125    QC lift to dense H (kept for API parity; the
                                                decoders never materialise H)
  create_LLR_mapping    ldpc_utils.py:62-95     variable-major edge numbering + neighbour tables
  get_LLR_indexes       ldpc_utils.py:5-60

`QCCode` holds the 197-entry (for BG2) shift table, uploads it to the device's constant
memory through ldpc_code_create, and can also be recovered from a dense H (so
`MinSumScaledDecoder(H, ...)` keeps working): every Z x Z block must be empty or a cyclic
permutation with check r connected to variable (r + s) mod Z (ldpc_utils.py:121-123).
Any binary H factors with Z = 1.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _native

_CODES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "codes")


# --------------------------------------------------------------------------------------
# reference-named helpers
# --------------------------------------------------------------------------------------
def load_base_matrix(file_path):
    """Whitespace-separated shift table -> float tensor (ldpc_utils.py:127-146)."""
    with open(file_path, "r") as f:
        rows = [[float(x) for x in line.split()] for line in f.readlines()]
    return torch.tensor(rows)


def expand_base_matrix(base_matrix, Z):
    """Dense (rows*Z, cols*Z) float32 H with I_Z rolled by `shift` columns (ldpc_utils.py:97-125)."""
    base = np.asarray(torch.as_tensor(base_matrix).cpu().numpy()).astype(np.int64)
    rows, cols = base.shape
    H = np.zeros((rows, Z, cols, Z), dtype=np.float32)
    r = np.arange(Z)
    for i, j in zip(*np.nonzero(base != -1)):
        H[i, r, j, (r + base[i, j]) % Z] = 1.0
    return torch.from_numpy(H.reshape(rows * Z, cols * Z))


def _padded(lists, width):
    out = np.full((len(lists), max(width, 0)), -1, dtype=np.int64)
    for i, l in enumerate(lists):
        out[i, :len(l)] = l
    return out


def get_LLR_indexes(H_to_LLR_mapping_T):
    """Neighbour tables (ldpc_utils.py:5-60): for every LLR (edge) index the other edges of
    the same check (rows of the mapping) and of the same variable (columns), -1 padded."""
    m = torch.as_tensor(H_to_LLR_mapping_T).cpu().numpy()
    E = int((m >= 0).sum())
    chk = [[] for _ in range(E)]
    var = [[] for _ in range(E)]
    for row in m:
        ids = row[row != -1]
        for x in ids:
            chk[x] = [y for y in ids if y != x]
    for col in m.T:
        ids = col[col != -1]
        for x in ids:
            var[x] = [y for y in ids if y != x]
    c = _padded(chk, max(len(v) for v in chk))
    v = _padded(var, max(len(v) for v in var))
    return torch.from_numpy(c), torch.from_numpy(v)


def create_LLR_mapping(H_T):
    """(H_to_LLR_mapping_T, check_LLR_matrix, var_LLR_matrix, output_index_tensor) for the
    TRANSPOSED parity-check matrix, edges numbered variable-major (ldpc_utils.py:62-95)."""
    Ht = torch.as_tensor(H_T).cpu().numpy()
    rows, cols = np.nonzero(Ht == 1)                     # row-major over H_T = variable-major
    mapping = np.full(Ht.shape, -1, dtype=np.int64)
    mapping[rows, cols] = np.arange(rows.size)
    mapping_T = np.ascontiguousarray(mapping.T)
    c, v = get_LLR_indexes(mapping_T)
    return torch.from_numpy(mapping_T), c, v, torch.from_numpy(rows.astype(np.int64)).unsqueeze(0)


# --------------------------------------------------------------------------------------
# the engine's code object
# --------------------------------------------------------------------------------------
class QCCode:
    """Quasi-cyclic LDPC code as a base-graph shift table (host copy + per-device handles)."""

    MAX_Z = 32
    MAX_SPLIT_Z = 512           # largest lifting factor looked for in a dense H (5G NR: 384)

    def __init__(self, shifts, Z):
        shifts = np.asarray(shifts)
        if shifts.ndim != 2:
            raise ValueError("base graph must be 2-D")
        shifts = np.rint(shifts).astype(np.int64)
        Z = int(Z)
        if not 1 <= Z <= self.MAX_Z:
            raise ValueError(f"lifting factor Z={Z} outside 1..{self.MAX_Z}")
        if (shifts < -1).any() or (shifts >= Z).any():
            raise ValueError("shift values must be -1 (empty) or in [0, Z)")
        self.shifts = shifts.astype(np.int16)
        self.Z = Z
        self.rows, self.cols = self.shifts.shape
        self.N, self.M = self.cols * Z, self.rows * Z
        self.base_edges = int((self.shifts >= 0).sum())
        self.E = self.base_edges * Z
        self.K = self.N - self.M
        self._handles = {}
        # set by from_base_matrix(..., allow_split=True) for a lifting factor above MAX_Z: the code is held as an equivalent
        # QC code with a smaller lifting factor and its variables / checks renumbered (see _split_lifting)
        self.lift_Z = None              # the caller's lifting factor
        self.lift_shifts = None         # the caller's base graph (shifts mod lift_Z)
        self.var_old_of_new = None      # (N,) int64: caller's variable index of each engine variable
        self.var_new_of_old = None      # (N,) int64: the inverse

    # ---- constructors ----
    @classmethod
    def from_base_matrix(cls, base_matrix, Z, allow_split=False):
        """Base graph as the reference's files hold it: non-negative shifts are taken mod Z, exactly what
        expand_base_matrix's roll does (ldpc_utils.py:121-123), so `NR_2_0_32.txt` lifted with Z=16 works as it
        does upstream.  The raw constructor keeps the strict [0, Z) check.

        Z > 32 (the reference accepts any --lifting_factor, main.py:38): with allow_split the code is rewritten as an
        equivalent QC code with lifting factor Zs = the largest divisor of Z that is <= 32 (_split_lifting) and a
        renumbering of its variables; the classic decoders apply the renumbering to their inputs and outputs, so the
        caller sees the natural order.  Other consumers (GNN, encoder, simulation) do not take such codes."""
        s = np.rint(np.asarray(torch.as_tensor(base_matrix).cpu().numpy(), dtype=np.float64)).astype(np.int64)
        Z = int(Z)
        if Z <= cls.MAX_Z or not allow_split:
            return cls(np.where(s >= 0, s % Z, -1), Z)
        nat = np.where(s >= 0, s % Z, -1)
        Zs = max(z for z in range(1, cls.MAX_Z + 1) if Z % z == 0)
        code = cls(cls._split_lifting(nat, Z, Zs), Zs)
        m = Z // Zs
        new = np.arange(code.N, dtype=np.int64)
        jb, a = np.divmod(new, Zs)
        j, b = np.divmod(jb, m)
        code.var_old_of_new = j * Z + m * a + b
        code.var_new_of_old = np.empty_like(code.var_old_of_new)
        code.var_new_of_old[code.var_old_of_new] = new
        code.lift_Z, code.lift_shifts = Z, nat.astype(np.int64)
        return code

    @staticmethod
    def _split_lifting(shifts, Z, Zs):
        """A Z x Z circulant with Z = m * Zs becomes an m x m block pattern of Zs x Zs circulants once rows and columns
        are renumbered r = m * a + b  ->  (b, a): row (b, a) of the shift-s circulant (s = m * q + t) meets column
        (b', a') with b' = (b + t) mod m and a' = a + q + carry, carry = (b + t) div m, i.e. block (b, b') is the
        circulant with shift (q + carry) mod Zs.  Base row i / column j and, inside them, the order of the cells are
        kept, so every check still meets its variables in ascending order and every variable its checks: the decoders'
        operation order -- and with it every bit of their output -- is that of the natural numbering."""
        m = Z // Zs
        rows, cols = shifts.shape
        out = np.full((rows * m, cols * m), -1, dtype=np.int64)
        for i, j in zip(*np.nonzero(shifts >= 0)):
            q, t = divmod(int(shifts[i, j]), m)
            for b in range(m):
                out[i * m + b, j * m + (b + t) % m] = (q + (b + t) // m) % Zs
        return out

    @classmethod
    def from_file(cls, path, Z):
        return cls.from_base_matrix(load_base_matrix(path), Z)

    @classmethod
    def nr_2_0(cls, Z):
        """The shipped 5G NR base graph 2, set index 0 (reference `NR_2_0_<Z>.txt`)."""
        if Z not in (2, 4, 8, 16, 32):
            raise ValueError("set index 0 of BG2 has lifting sizes 2, 4, 8, 16, 32 (within Z <= 32)")
        rows = cols = None
        cells = []
        with open(os.path.join(_CODES_DIR, "bg2_ils0_mod32.triples")) as f:
            for line in f:
                tok = line.split()
                if not tok or tok[0].startswith("#"):
                    continue
                if tok[0] == "rows":
                    rows, cols = int(tok[1]), int(tok[3])
                else:
                    cells.append((int(tok[0]), int(tok[1]), int(tok[2])))
        s = np.full((rows, cols), -1, dtype=np.int64)
        for i, j, v in cells:
            s[i, j] = v % Z
        return cls(s, Z)

    @classmethod
    def from_dense(cls, H, Z=None, allow_split=False):
        """Factor a dense 0/1 parity-check matrix into circulant shifts.  Tries the given Z,
        else the largest Z <= 32 dividing both dimensions for which every block is empty or a
        cyclic permutation; Z = 1 always succeeds.  Raises ValueError if `Z` was given and
        H is not quasi-cyclic with that lifting factor.

        allow_split (classic decoders): lifting factors above 32 are looked for first, largest first -- the reference's
        scripts hand the decoders `expand_base_matrix(base, --lifting_factor)` (main.py:92,152,215) -- and held as an
        equivalent renumbered code (from_base_matrix)."""
        Hn = torch.as_tensor(H).detach().cpu().numpy()
        if Hn.ndim != 2:
            raise ValueError("H must be 2-D")
        if not np.isin(Hn, (0, 1)).all():
            raise ValueError("H must be binary")
        Hb = Hn.astype(np.uint8)
        M, N = Hb.shape
        if allow_split and (Z is None or int(Z) > cls.MAX_Z):
            g = int(np.gcd(M, N))
            big = [int(Z)] if Z is not None else [z for z in range(min(g, cls.MAX_SPLIT_Z), cls.MAX_Z, -1) if g % z == 0]
            for z in big:
                if M % z or N % z:
                    raise ValueError(f"H of shape {Hb.shape} cannot be lifted with Z={z}")
                s = cls._factor(Hb, z)
                if s is not None:
                    return cls.from_base_matrix(s, z, allow_split=True)
            if Z is not None:
                raise ValueError(f"H is not quasi-cyclic with Z={Z}")
        cand = [int(Z)] if Z is not None else [z for z in range(cls.MAX_Z, 0, -1) if M % z == 0 and N % z == 0]
        for z in cand:
            if z < 1 or z > cls.MAX_Z or M % z or N % z:
                raise ValueError(f"H of shape {Hb.shape} cannot be lifted with Z={z}")
            s = cls._factor(Hb, z)
            if s is not None:
                return cls(s, z)
        raise ValueError(f"H is not quasi-cyclic with Z={Z}")

    @staticmethod
    def _factor(Hb, Z):
        M, N = Hb.shape
        rows, cols = M // Z, N // Z
        blk = Hb.reshape(rows, Z, cols, Z).transpose(0, 2, 1, 3)          # (rows, cols, Z, Z)
        nnz = blk.sum(axis=(2, 3))
        if not np.isin(nnz, (0, Z)).all():
            return None
        first = blk[:, :, 0, :].argmax(axis=2)                            # shift candidate from block row 0
        r = np.arange(Z)
        want = np.zeros((rows, cols, Z, Z), dtype=np.uint8)
        ii, jj = np.nonzero(nnz == Z)
        for i, j in zip(ii, jj):
            want[i, j, r, (r + first[i, j]) % Z] = 1
        if not np.array_equal(want, blk):
            return None
        return np.where(nnz == Z, first, -1)

    # ---- views ----
    def base_matrix(self):
        """The base graph in the caller's terms (for a split code: the one it was built from, to be lifted with lift_Z)."""
        src = self.shifts if self.lift_shifts is None else self.lift_shifts
        return torch.from_numpy(src.astype(np.float32))

    def dense(self):
        return expand_base_matrix(self.base_matrix(), self.Z if self.lift_Z is None else self.lift_Z)

    def edges(self):
        """(check, variable) arrays of all E Tanner edges in check-major, ascending-variable
        order (the order of the reference's adjacency lists and GNN message list)."""
        chk, var = [], []
        r = np.arange(self.Z)
        for i in range(self.rows):
            per = []
            for j in range(self.cols):
                s = self.shifts[i, j]
                if s >= 0:
                    per.append((i * self.Z + r, j * self.Z + (r + s) % self.Z))
            if per:
                c = np.stack([p[0] for p in per], axis=1).reshape(-1)     # row r: its edges in ascending j
                v = np.stack([p[1] for p in per], axis=1).reshape(-1)
                chk.append(c)
                var.append(v)
        return np.concatenate(chk), np.concatenate(var)

    # ---- device handle ----
    def handle(self, device):
        """Native handle (tables uploaded to `device`), created once per device."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the LDPC engine runs on CUDA devices only (no CPU fallback)")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(index)
        if h is None:
            flat = np.ascontiguousarray(self.shifts.reshape(-1))
            out = C.c_void_p()
            _native.check(_native.lib().ldpc_code_create(
                flat.ctypes.data_as(C.c_void_p), self.rows, self.cols, self.Z, index, C.byref(out)))
            h = _Handle(out)
            self._handles[index] = h
        return h.ptr

    def has_fast_path(self, device, algo=_native.ALGO_MINSUM):
        return bool(_native.lib().ldpc_code_has_fast_path(self.handle(device), algo))

    def __repr__(self):
        return f"QCCode(rows={self.rows}, cols={self.cols}, Z={self.Z}, N={self.N}, M={self.M}, E={self.E})"


class _Handle:
    def __init__(self, p):
        self.ptr = p

    def __del__(self):
        try:
            _native.lib().ldpc_code_destroy(self.ptr)
        except Exception:
            pass


def as_code(H=None, base_graph=None, Z=None, allow_split=False):
    """Resolve the constructor arguments of the drop-in decoders to a QCCode (allow_split: see QCCode.from_base_matrix)."""
    if isinstance(H, QCCode):
        if H.lift_Z is not None and not allow_split:
            raise ValueError(f"a code with lifting factor {H.lift_Z} > {QCCode.MAX_Z} is supported by the classic decoders only")
        return H
    if base_graph is not None:
        if Z is None:
            raise ValueError("base_graph given without Z")
        return QCCode.from_base_matrix(base_graph, Z, allow_split=allow_split)
    if H is None:
        raise ValueError("need a parity-check matrix H, a QCCode, or (base_graph, Z)")
    return QCCode.from_dense(H, Z, allow_split=allow_split)
