// neural.cuh -- the unrolled neural min-sum decoder (LDPCNeuralDecoder) as ONE kernel.
//
// Replaces, for inference / validation, the per-iteration chain CheckLayer -> VariableLayer ->
// ResidualLayer and the final OutputLayer of models/layers.py (:14-66, :78-125, :143-168,
// :180-210) as composed by the reference's missing models/decoder.py (prototype:
// EE4002R_2025.ipynb cell 11 `forward`).  Every arithmetic step is the one of the per-layer
// kernels in layers.cuh, in the same order, so the output is bit-identical to the composition;
// what changes is where the state lives:
//   * a CTA keeps kRows codewords resident in shared memory for ALL iterations: the channel LLRs,
//     the check messages c2v[kRows][E] and a ring of max(L,1) earlier variable outputs x[kRows][E]
//     (x_l doubles as the input of the next check layer and as prev[0] of the next residual;
//     the new x overwrites the oldest ring slot in place -- each thread reads and writes only its
//     own element there).  HBM sees one coalesced read of llr_e and one coalesced write of the soft
//     outputs per codeword (25 KB each at BG2 Z=32) instead of six [B,E] round trips per iteration.
//   * the neighbour tables are pre-packed once per code to uint16, k-major ([K][E], 0xFFFF = the
//     reference's -1 padding): a warp's index load is one coalesced 64 B segment instead of 32
//     strided 8-byte words, 4x fewer bytes, and L2-resident (390 KB at Z=32), shared by the
//     kRows codewords of the CTA.
// Algorithmic traffic per codeword: 4E in + 4E soft out (+ 4E ground truth) = 50 KB at Z=32;
// the kernel is bound by shared-memory gathers (sum_d d(d-1) = 93 376 per codeword-iteration).
#pragma once
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "layers.cuh"

namespace ldpc {

constexpr int kNeuralThreads = 1024;
constexpr unsigned short kNeuralPad = 0xFFFFu;
#ifndef LDPC_NEURAL_GROUP
#define LDPC_NEURAL_GROUP 4
#endif
constexpr int kNeuralGroup = LDPC_NEURAL_GROUP;

// idx [E,K] int64 (-1 padded) -> out [K,E] uint16 (0xFFFF padded)
__global__ void neural_pack_index_kernel(const long long* __restrict__ idx, long long E, int K,
                                         unsigned short* __restrict__ out) {
    const long long total = E * K;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long k = t / E, e = t - k * E;
        const long long n = idx[e * K + k];
        out[t] = n < 0 ? kNeuralPad : (unsigned short)n;
    }
}

// KC / KV: compile-time neighbour-table widths (0 = use the run-time Kc / Kv).  With the widths of
// the 5G BG2 tables (9 / 22) known, the slot loops unroll completely: a thread has all its index
// loads in flight before the first shared-memory gather, and the loop stops at the warp's last
// used slot.  cperm / vperm (optional): thread t works on edge perm[t] and reads column t of the
// table, which the host has sorted by descending neighbour count (padding compacted to the end of
// each row) -- warps then see uniform list lengths and skip the padded slots, 59 % of the check
// table and 55 % of the variable table at BG2.  Results do not depend on the permutation: per edge,
// the valid neighbours are visited in the caller's order.

// sum over the valid slots of column t of src[q][idx]; all 32 lanes of the warp must call it
template <int kRows, int KV>
__device__ __forceinline__ void neural_gather_sum(const float* __restrict__ src, const unsigned short* __restrict__ idx,
                                                  int Kv, int E, int t, bool live, float (&acc)[kRows]) {
#pragma unroll
    for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
    if constexpr (KV > 0) {
        unsigned short nn[KV];
        int last = 0;
#pragma unroll
        for (int k = 0; k < KV; ++k) {
            nn[k] = live ? idx[k * E + t] : kNeuralPad;
            last = nn[k] != kNeuralPad ? k + 1 : last;
        }
        const int kmax = __reduce_max_sync(0xffffffffu, last);
        // groups of kNeuralGroup slots between warp-uniform exits: inside a group the gathers are independent and
        // predicated, so they are in flight together
#pragma unroll
        for (int k0 = 0; k0 < KV; k0 += kNeuralGroup) {
            if (k0 >= kmax) break;
            float v[kNeuralGroup][kRows];
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j) {
                const bool on = k0 + j < KV && nn[k0 + j < KV ? k0 + j : 0] != kNeuralPad;
                const int n = on ? nn[k0 + j < KV ? k0 + j : 0] : 0;
#pragma unroll
                for (int q = 0; q < kRows; ++q) v[j][q] = on ? src[q * E + n] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j) {
                const bool on = k0 + j < KV && nn[k0 + j < KV ? k0 + j : 0] != kNeuralPad;
#pragma unroll
                for (int q = 0; q < kRows; ++q) acc[q] = on ? acc[q] + v[j][q] : acc[q];
            }
        }
    } else {
        for (int k = 0; k < Kv; ++k) {
            const unsigned short n = live ? idx[k * E + t] : kNeuralPad;
            if (n == kNeuralPad) continue;
#pragma unroll
            for (int q = 0; q < kRows; ++q) acc[q] += src[q * E + n];
        }
    }
}

// CheckLayer arithmetic for one valid neighbour value (layers.cuh check_layer_fwd_kernel): the product of
// torch.sign(v + 1e-10) factors is kept as (xor of sign bits, "a factor was 0") -- the same +-1 / +-0 result
// as the sequential float product -- and zeros count as 1e10 in the minimum.
__device__ __forceinline__ void neural_check_visit(float v, unsigned& negb, bool& zero, float& mn) {
    const float sh = __fadd_rn(v, 1e-10f);
    negb ^= __float_as_uint(sh);
    zero |= !(fabsf(sh) > 0.0f);                         // 0 (or NaN, which torch-style comparisons also map to 0)
    float a = fabsf(v);
    a = a != 0.0f ? a : 1e10f;
    mn = a < mn ? a : mn;
}

template <int kRows, int KC, int KV>
__global__ void __launch_bounds__(kNeuralThreads) neural_decode_kernel(
    const float* __restrict__ llr, const unsigned short* __restrict__ cidx, int Kc_rt,
    const unsigned short* __restrict__ cperm, const unsigned short* __restrict__ vidx, int Kv_rt,
    const unsigned short* __restrict__ vperm, const float* __restrict__ w_ch, const float* __restrict__ w_res, int L,
    int iters, long long B, int E, const float* __restrict__ gt, float* __restrict__ soft,
    float* __restrict__ max_loss) {
    extern __shared__ float sm[];
    const int Kc = KC ? KC : Kc_rt, Kv = KV ? KV : Kv_rt;
    const int Lb = L > 0 ? L : 1;
    const int RE = kRows * E;
    float* c2v = sm;                                  // [kRows][E] check-to-variable messages
    float* lls = sm + RE;                             // [kRows][E] channel LLRs of the resident codewords
    float* ring = sm + 2 * RE;                        // [Lb][kRows][E] earlier variable outputs
    __shared__ float red[kRows][kNeuralThreads / 32];
    const int lane = threadIdx.x & 31;

    for (long long b0 = (long long)blockIdx.x * kRows; b0 < B; b0 += (long long)gridDim.x * kRows) {
        const int nb = (int)((B - b0) < kRows ? (B - b0) : kRows);
        __syncthreads();
        // x_0 = llr_e sits in ring slot 0 but is NOT a queue entry (notebook cell 11: no residual in the first
        // update); rows past the end of the batch repeat the last row (their results are never stored)
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
            const long long row = b0 + (q < nb ? q : nb - 1);
            for (int e = threadIdx.x; e < E; e += kNeuralThreads) {
                const float v = llr[row * E + e];
                lls[q * E + e] = v;
                ring[q * E + e] = v;
            }
        }
        __syncthreads();
        int cur = 0, nq = 0;                          // newest ring slot, queue length
        for (int l = 0; l < iters; ++l) {
            // ---- CheckLayer on x = ring[cur] ----
            const float* x = ring + cur * RE;
            for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
                const bool live = t < E;
                const int tt = live ? t : 0;
                const int e = cperm ? cperm[tt] : tt;
                float mn[kRows];
                unsigned negb[kRows];
                bool zero[kRows];
#pragma unroll
                for (int q = 0; q < kRows; ++q) { mn[q] = CUDART_INF_F; negb[q] = 0u; zero[q] = false; }
                int used = 0;
                if constexpr (KC > 0) {
                    unsigned short nn[KC];
                    int last = 0;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        nn[k] = live ? cidx[k * E + tt] : kNeuralPad;
                        last = nn[k] != kNeuralPad ? k + 1 : last;
                        used += nn[k] != kNeuralPad;
                    }
                    const int kmax = __reduce_max_sync(0xffffffffu, last);
#pragma unroll
                    for (int k0 = 0; k0 < KC; k0 += kNeuralGroup) {
                        if (k0 >= kmax) break;           // warp-uniform
                        float v[kNeuralGroup][kRows];
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j) {
                            const bool on = k0 + j < KC && nn[k0 + j < KC ? k0 + j : 0] != kNeuralPad;
                            const int n = on ? nn[k0 + j < KC ? k0 + j : 0] : 0;
#pragma unroll
                            for (int q = 0; q < kRows; ++q) v[j][q] = x[q * E + n];
                        }
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j) {
                            const bool on = k0 + j < KC && nn[k0 + j < KC ? k0 + j : 0] != kNeuralPad;
                            if (on) {
#pragma unroll
                                for (int q = 0; q < kRows; ++q) neural_check_visit(v[j][q], negb[q], zero[q], mn[q]);
                            }
                        }
                    }
                } else {
                    for (int k = 0; k < Kc; ++k) {
                        const unsigned short n = live ? cidx[k * E + tt] : kNeuralPad;
                        if (n == kNeuralPad) continue;
                        ++used;
#pragma unroll
                        for (int q = 0; q < kRows; ++q) neural_check_visit(x[q * E + n], negb[q], zero[q], mn[q]);
                    }
                }
                if (live) {
#pragma unroll
                    for (int q = 0; q < kRows; ++q) {
                        // a padded slot is a zero input: sign factor +1, magnitude 1e10 (layers.py:48-57)
                        const float m = (used < Kc && 1e10f < mn[q]) ? 1e10f : mn[q];
                        const float sp = __uint_as_float((negb[q] & 0x80000000u) | (zero[q] ? 0u : 0x3f800000u));
                        c2v[q * E + e] = sp * m;
                    }
                }
            }
            __syncthreads();
            if (l == iters - 1) break;
            // ---- VariableLayer(0, c2v) + ResidualLayer (layers.cuh neural_variable_fwd_kernel) ----
            const int nxt = nq == 0 ? 0 : (cur + 1) % Lb;   // empty slot, or the oldest entry once the queue is full
            for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
                const bool live = t < E;
                const int tt = live ? t : 0;
                const int e = vperm ? vperm[tt] : tt;
                float acc[kRows];
                neural_gather_sum<kRows, KV>(c2v, vidx, Kv, E, tt, live, acc);
                if (live) {
                    const float w = w_ch[e];
#pragma unroll
                    for (int q = 0; q < kRows; ++q) {
                        float r = __fadd_rn(__fmul_rn(lls[q * E + e], w), acc[q]);
                        for (int i = 0; i < nq; ++i) {
                            const int slot = (cur - i + Lb) % Lb;
                            r = __fadd_rn(r, __fmul_rn(__ldg(w_res + i), ring[slot * RE + q * E + e]));
                        }
                        ring[nxt * RE + q * E + e] = r;
                    }
                }
            }
            cur = nxt;
            nq = nq < L ? nq + 1 : nq;
            __syncthreads();
        }
        // ---- final = VariableLayer(c2v, c2v); OutputLayer(final, llr, gt).  The ring is free now: soft values
        //      are staged there so that the global stores (and the ground-truth loads) are coalesced. ----
        float* stage = ring;
        for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
            const bool live = t < E;
            const int tt = live ? t : 0;
            const int e = vperm ? vperm[tt] : tt;
            float acc[kRows];
            neural_gather_sum<kRows, KV>(c2v, vidx, Kv, E, tt, live, acc);
            if (live) {
#pragma unroll
                for (int q = 0; q < kRows; ++q) {
                    const float fin = __fadd_rn(c2v[q * E + e], acc[q]);
                    const float z = __fadd_rn(fin, lls[q * E + e]);
                    stage[q * E + e] = 1.0f / (1.0f + expf(-z));
                }
            }
        }
        __syncthreads();
        float best[kRows];
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
            best[q] = -CUDART_INF_F;
            if (q < nb) {
                for (int e = threadIdx.x; e < E; e += kNeuralThreads) {
                    const long long g = (b0 + q) * E + e;
                    const float s = stage[q * E + e];
                    soft[g] = s;
                    if (gt) {
                        const float y = gt[g];
                        const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                        const float loss = -(y * l1 + (1.0f - y) * l0);
                        best[q] = loss > best[q] ? loss : best[q];
                    }
                }
            }
        }
        if (gt) {
#pragma unroll
            for (int q = 0; q < kRows; ++q) {
                float v = best[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) red[q][threadIdx.x >> 5] = v;
            }
            __syncthreads();
            if (threadIdx.x < 32) {
#pragma unroll
                for (int q = 0; q < kRows; ++q) {
                    float v = red[q][threadIdx.x];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (threadIdx.x == 0 && q < nb) max_loss[b0 + q] = v;
                }
            }
        }
    }
}

}  // namespace ldpc
