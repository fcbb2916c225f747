// decode_fast.cuh -- host-side selection of the specialised kernels (decode_fast_kernel.cuh).
// The kernels are compiled in their own translation unit (fast_kernels.cu).
#pragma once
#include "bg2_tables.h"
#include "params.cuh"
#include "tables.cuh"

namespace ldpc {

template <class BG>
inline bool table_matches(const ldpc_code* c) {
    if (c->rows != BG::kRows || c->cols != BG::kCols || c->Z != BG::kZ || c->E != BG::kEdges) return false;
    const uint32_t* t = c->h_tab.data();
    const int off_rowptr = t[7], off_redge = t[9];
    for (int i = 0; i <= BG::kRows; ++i)
        if ((int)t[off_rowptr + i] != BG::row_ptr[i]) return false;
    for (int e = 0; e < BG::kEdges; ++e) {
        const uint32_t w = t[off_redge + e];
        if ((int)(w & 0xffff) != BG::col[e] || (int)((w >> 16) & 0xff) != BG::shift[e]) return false;
    }
    return true;
}

inline int detect_fast_kind(const ldpc_code* c) {
    if (table_matches<BG2Z32>(c)) return 1;
    if (table_matches<BG2Z4>(c)) return 2;
    if (table_matches<BG2Z16>(c)) return 3;
    if (table_matches<BG2Z8>(c)) return 4;
    return 0;
}

// fixed iteration count: every compiled table; per-codeword early exit: min-sum on every table, sum-product at Z = 32 and 16;
// hard decisions / syndrome / iteration counts only (no soft output, no validity masks -- those stay on the exact kernel)
inline bool fast_path_supports(const ldpc_code* c, int algo, int stop_mode, bool want_mask, bool want_soft) {
    if (c->fast_kind == 0 || (algo != LDPC_ALGO_MINSUM && algo != LDPC_ALGO_BP) || want_mask) return false;
    if (stop_mode == LDPC_STOP_FIXED) return true;
    if (stop_mode != LDPC_STOP_PER_CODEWORD || want_soft) return false;
    return algo == LDPC_ALGO_MINSUM || c->fast_kind == 1 || c->fast_kind == 3;
}

// defined in fast_kernels.cu
int launch_fast(const ldpc_code* c, int algo, const DecodeParams& p, cudaStream_t st);

}  // namespace ldpc
