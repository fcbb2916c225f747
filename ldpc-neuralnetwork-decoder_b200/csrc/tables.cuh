// tables.cuh -- quasi-cyclic code description resident on the device.
//
// Replaces the reference's dense H (utils/ldpc_utils.py:97-125) and the Python adjacency
// lists built from it (models/traditional_decoders.py:26-40).  A code is `rows x cols` base
// cells, each either empty or a Z x Z circulant with shift s: check i*Z+r is connected to
// variable j*Z+((r+s) mod Z).  Any binary H is QC with Z=1, so the same tables (and the same
// kernels) also serve non-QC matrices such as the notebook's 3x4 toy H.
//
// Flat word layout (uint32), used both in __constant__ slots and in global memory:
//   [0] rows [1] cols [2] Z [3] E(base edges) [4] G = codewords per warp = 32/Z
//   [5] max row degree [6] max col degree [7] off_rowptr [8] off_colptr [9] off_redge
//   [10] off_cedge [11] total words
//   rowptr[rows+1], colptr[cols+1]
//   redge[E]  row-major, ascending column:  col | shift<<16 | (col degree==1)<<24
//   cedge[E]  col-major, ascending row:     edge id | shift<<16
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ldpc {

constexpr int kTabHdr = 16;
constexpr int kSlotWords = 2048;   // 8 KB per slot: BG2 needs 506 words, BG1 (46x68, 316 cells) 764
constexpr int kNumSlots = 6;       // 48 KB of the 64 KB constant bank
constexpr int kMaxDevices = 16;

__constant__ uint32_t c_tab[kNumSlots][kSlotWords];

// Table accessor: LDC through the constant bank when the code owns a slot, LDG otherwise.
template <bool kConst>
struct Tab {
    const uint32_t* g;
    int slot;
    __device__ __forceinline__ uint32_t operator[](int i) const {
        if constexpr (kConst) return c_tab[slot][i];
        else return __ldg(g + i);
    }
};

}  // namespace ldpc

// Device staging buffers + streams of ldpc_decode_host, owned by the handle and reused
// across calls (H2D / decode / D2H of consecutive chunks overlap on three streams).
struct HostStage {
    static constexpr int kStages = 3;
    std::mutex mu;
    cudaStream_t st[kStages] = {};
    void* d_llr[kStages] = {};
    void* d_raw[kStages] = {};     // quantised LLRs as transferred (ldpc_decode_host_q)
    void* d_soft[kStages] = {};
    char* d_hard[kStages] = {};
    size_t cap_llr[kStages] = {}, cap_soft[kStages] = {}, cap_hard[kStages] = {}, cap_raw[kStages] = {};
};

struct ldpc_code {
    int device = 0;
    int rows = 0, cols = 0, Z = 0, E = 0, N = 0, M = 0, G = 0, maxdc = 0, maxdv = 0;
    int slot = -1;                 // constant slot, -1 = global only
    uint32_t* d_tab = nullptr;     // global copy (always present)
    int tab_words = 0;
    int fast_kind = 0;             // 0 none; 5G BG2 set 0 at Z = 32 (1, the NR_2_0_32 table), 4 (2), 16 (3), 8 (4)
    std::vector<uint32_t> h_tab;
    std::vector<int16_t> shifts;   // rows*cols
    mutable HostStage stage;
};

namespace ldpc {

inline std::mutex& slot_mutex() {
    static std::mutex m;
    return m;
}
inline bool (&slot_used())[kMaxDevices][kNumSlots] {
    static bool used[kMaxDevices][kNumSlots] = {};
    return used;
}

// Build the flat table on the host.  Returns false (with message) on malformed input.
inline int build_table(const int16_t* shifts, int rows, int cols, int Z, ldpc_code* c) {
    if (!shifts || rows <= 0 || cols <= 0) return fail(LDPC_ERR_INVALID, "code_create: null shifts or empty base graph");
    if (Z < 1 || Z > 32) return fail(LDPC_ERR_UNSUPPORTED, "code_create: lifting factor Z=%d outside 1..32", Z);
    if (cols >= 65536) return fail(LDPC_ERR_UNSUPPORTED, "code_create: more than 65535 base columns");
    std::vector<int> rowdeg(rows, 0), coldeg(cols, 0);
    int E = 0;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            int s = shifts[(size_t)i * cols + j];
            if (s < -1 || s >= Z) return fail(LDPC_ERR_INVALID, "code_create: shift %d at (%d,%d) not in [-1,%d)", s, i, j, Z);
            if (s >= 0) { ++rowdeg[i]; ++coldeg[j]; ++E; }
        }
    if (E == 0) return fail(LDPC_ERR_INVALID, "code_create: base graph has no edges");
    if (E >= 65536) return fail(LDPC_ERR_UNSUPPORTED, "code_create: more than 65535 base edges");
    c->rows = rows; c->cols = cols; c->Z = Z; c->E = E; c->N = cols * Z; c->M = rows * Z; c->G = 32 / Z;
    c->maxdc = 0; c->maxdv = 0;
    for (int d : rowdeg) c->maxdc = d > c->maxdc ? d : c->maxdc;
    for (int d : coldeg) c->maxdv = d > c->maxdv ? d : c->maxdv;
    const int off_rowptr = kTabHdr, off_colptr = off_rowptr + rows + 1, off_redge = off_colptr + cols + 1,
              off_cedge = off_redge + E, total = off_cedge + E;
    std::vector<uint32_t>& t = c->h_tab;
    t.assign(total, 0);
    t[0] = rows; t[1] = cols; t[2] = Z; t[3] = E; t[4] = c->G; t[5] = c->maxdc; t[6] = c->maxdv;
    t[7] = off_rowptr; t[8] = off_colptr; t[9] = off_redge; t[10] = off_cedge; t[11] = total;
    // row-major edge list
    std::vector<int> edge_of(rows * cols, -1);
    int e = 0;
    for (int i = 0; i < rows; ++i) {
        t[off_rowptr + i] = e;
        for (int j = 0; j < cols; ++j) {
            int s = shifts[(size_t)i * cols + j];
            if (s < 0) continue;
            t[off_redge + e] = (uint32_t)j | ((uint32_t)s << 16) | ((coldeg[j] == 1 ? 1u : 0u) << 24);
            edge_of[i * cols + j] = e++;
        }
    }
    t[off_rowptr + rows] = e;
    int k = 0;
    for (int j = 0; j < cols; ++j) {
        t[off_colptr + j] = k;
        for (int i = 0; i < rows; ++i) {
            int s = shifts[(size_t)i * cols + j];
            if (s < 0) continue;
            t[off_cedge + k++] = (uint32_t)edge_of[i * cols + j] | ((uint32_t)s << 16);
        }
    }
    t[off_colptr + cols] = k;
    c->tab_words = total;
    c->shifts.assign(shifts, shifts + (size_t)rows * cols);
    return LDPC_OK;
}

}  // namespace ldpc
