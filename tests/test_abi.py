"""The C-ABI library loads and exports every symbol include/ldpc_b200.h declares; error paths
that need no GPU behave as documented.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import torch

from conftest import ROOT
import ldpc_b200
from ldpc_b200 import _native


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 25
    handle = C.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
        assert n in _native.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_native.PROTOTYPES) <= set(names)
    assert _native.lib().ldpc_abi_version() == 1


def test_argument_errors_without_gpu():
    L = _native.lib()
    out = C.c_void_p()
    shifts = np.array([0, -1, 1, 0], dtype=np.int16)
    assert L.ldpc_code_create(None, 2, 2, 4, 0, C.byref(out)) == _native.ERR_INVALID
    assert b"null" in L.ldpc_last_error()
    assert L.ldpc_code_create(shifts.ctypes.data_as(C.c_void_p), 2, 2, 64, 0, C.byref(out)) == _native.ERR_UNSUPPORTED
    bad = np.array([0, 9, 1, 0], dtype=np.int16)
    assert L.ldpc_code_create(bad.ctypes.data_as(C.c_void_p), 2, 2, 4, 0, C.byref(out)) == _native.ERR_INVALID
    assert L.ldpc_minsum_decode(None, None, 1, 1, 0.75, 0, 0, None, None, 0, None, None, None, 0, None) == _native.ERR_INVALID
    assert L.ldpc_sim_fer(None, 0, 1, 0.75, 0.0, 0, 0, 1, None, None) == _native.ERR_INVALID
    if not torch.cuda.is_available():
        # no device: creating a code must fail loudly, not fall back
        rc = L.ldpc_code_create(shifts.ctypes.data_as(C.c_void_p), 2, 2, 4, 0, C.byref(out))
        assert rc == _native.ERR_CUDA and not out.value
