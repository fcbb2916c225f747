"""ctypes binding of the C ABI in include/ldpc_b200.h (libldpc_b200.so, built in-tree).

There is no fallback of any kind: if the shared library is missing, `lib()` raises; if no
CUDA device is present, the first compute call returns LDPC_ERR_CUDA and `check()` raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDPC_B200_LIB: load an experimental build (tools/build_variant.sh) instead of the in-tree library
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(_HERE, "libldpc_b200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4
HARD_F32, HARD_U8, HARD_PACKED = 0, 1, 2
STOP_FIXED, STOP_PER_CODEWORD = 0, 1
PATH_AUTO, PATH_EXACT, PATH_FAST = 0, 1, 2
ALGO_MINSUM, ALGO_BP = 0, 1
LLR_F32, LLR_F16, LLR_I8 = 0, 1, 2
PATHS = {"auto": PATH_AUTO, "exact": PATH_EXACT, "fast": PATH_FAST}

_p, _i, _i64, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/ldpc_b200.h declaration by declaration
PROTOTYPES = {
    "ldpc_abi_version": (_i, []),
    "ldpc_last_error": (C.c_char_p, []),
    "ldpc_launch_count": (_u64, []),
    "ldpc_code_create": (_i, [_p, _i, _i, _i, _i, C.POINTER(_p)]),
    "ldpc_code_destroy": (_i, [_p]),
    "ldpc_code_info": (_i, [_p, C.POINTER(C.c_int32)]),
    "ldpc_code_has_fast_path": (_i, [_p, _i]),
    "ldpc_minsum_decode": (_i, [_p, _p, _i64, _i, _f, _i, _i, _p, _p, _i, _p, _p, _p, _i, _p]),
    "ldpc_bp_decode": (_i, [_p, _p, _i64, _i, _i, _i, _p, _p, _i, _p, _p, _p, _i, _p]),
    "ldpc_nonfinite_flag": (_i, [_p, _i64, _p, _p]),
    "ldpc_syndrome_check": (_i, [_p, _p, _i, _i64, _p, _p]),
    "ldpc_encode": (_i, [_p, _p, _i64, _p, _i64, _p, _p, _p]),
    "ldpc_rate_match": (_i, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "ldpc_rate_recover": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _p, _p]),
    "ldpc_decode_host": (_i, [_p, _i, _p, _i64, _i, _f, _i, _p, _p, _i, _i64]),
    "ldpc_decode_host_q": (_i, [_p, _i, _p, _i, _f, _i64, _i, _f, _i, _p, _p, _i, _i64]),
    "ldpc_awgn_llr": (_i, [_p, _i64, _i64, _f, _u64, _u64, _p, _p]),
    "ldpc_qpsk_llr": (_i, [_p, _i64, _i64, _f, _i, _u64, _u64, _p, _p]),
    "ldpc_count_errors": (_i, [_p, _i, _p, _i64, _i64, _p, _p]),
    "ldpc_sim_fer": (_i, [_p, _i, _i, _f, _f, _u64, _u64, _u64, _p, _p]),
    "ldpc_check_layer_fwd": (_i, [_p, _p, _i64, _i64, _i, _p, _p, _p]),
    "ldpc_check_layer_bwd": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _p, _p]),
    "ldpc_variable_layer_fwd": (_i, [_p, _p, _p, _i64, _i64, _i, _p, _p]),
    "ldpc_variable_layer_bwd": (_i, [_p, _p, _i64, _i64, _i, _p, _p]),
    "ldpc_neural_variable_layer_fwd": (_i, [_p, _p, _p, _p, _p, C.POINTER(_p), _i, _i64, _i64, _i, _p, _p]),
    "ldpc_neural_pack_index": (_i, [_p, _i64, _i, _p, _p]),
    "ldpc_neural_decode": (_i, [_p, _p, _i, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i64, _i64, _p, _p, _p, _p]),
    "ldpc_neural_decode_qc": (_i, [_p, _p, _p, _p, _i, _i, _i64, _p, _p, _p, _p, _p, _p]),
    "ldpc_neural_backward_qc": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _p]),
    "ldpc_neural_decode_qc_var": (_i, [_p, _p, _p, _p, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "ldpc_neural_backward_qc_var": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _p]),
    "ldpc_check_layer_fwd_sorted": (_i, [_p, _p, _i, _p, _p, _i64, _i64, _p, _p, _p]),
    "ldpc_variable_layer_fwd_sorted": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, C.POINTER(_p), _i, _i64, _i64, _p, _p]),
    "ldpc_check_layer_bwd_nstar": (_i, [_p, _p, _p, _p, _i64, _i64, _p, _p]),
    "ldpc_residual_layer_fwd": (_i, [_p, _p, _p, _p, C.POINTER(_p), _i, _i64, _i64, _p, _p]),
    "ldpc_output_layer_fwd": (_i, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "ldpc_gnn_create": (_i, [_p, _i, _i, _i, _p, C.POINTER(_p)]),
    "ldpc_gnn_destroy": (_i, [_p]),
    "ldpc_gnn_param_count": (_sz, [_p]),
    "ldpc_gnn_workspace_bytes": (_sz, [_p, _i64, _i]),
    "ldpc_gnn_forward": (_i, [_p, _p, _p, _i64, _p, _p, _p, _sz, _i, _p]),
    "ldpc_gnn_backward": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _sz, _p]),
}

_lib = None


class LdpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ldpc_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Load libldpc_b200.so once.  Raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA engine has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` at the repository root. "
                "There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.ldpc_abi_version() != 1:
            raise ImportError("libldpc_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc):
    if rc != OK:
        raise LdpcError(rc, lib().ldpc_last_error().decode("utf-8", "replace"))


def ptr(t):
    """data_ptr of a tensor or None -> c_void_p."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
