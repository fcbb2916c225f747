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
//   * the neighbour tables are pre-packed once per code to uint16, k-major ([K][E]): a warp's
//     index load is one coalesced 64 B segment instead of 32 strided 8-byte words, 4x fewer
//     bytes, and L2-resident (390 KB at Z=32), shared by the kRows codewords of the CTA.
// Algorithmic traffic per codeword: 4E in + 4E soft out (+ 4E ground truth) = 50 KB at Z=32;
// the kernel is bound by shared-memory gathers (sum_d d(d-1) = 93 376 per codeword-iteration).
#pragma once
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "layers.cuh"

namespace ldpc {

#ifndef LDPC_NEURAL_THREADS
#define LDPC_NEURAL_THREADS 1024
#endif
constexpr int kNeuralThreads = LDPC_NEURAL_THREADS;
constexpr unsigned short kNeuralPad = 0xFFFFu;
#ifndef LDPC_NEURAL_GROUP
#define LDPC_NEURAL_GROUP 8
#endif
constexpr int kNeuralGroup = LDPC_NEURAL_GROUP;
// 1: load a column's whole index list before the first gather.  Measured slower (2.3-2.6 M cw/s at 768 / 1024
// threads against 2.95 M with group-wise loads), kept for experiments.
#ifndef LDPC_NEURAL_PRELOAD
#define LDPC_NEURAL_PRELOAD 0
#endif
constexpr bool kNeuralPreload = LDPC_NEURAL_PRELOAD != 0;

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

// KC / KV: compile-time neighbour-table widths (0 = use the run-time Kc / Kv): with the widths of
// the 5G BG2 tables (9 / 22) the slot loops unroll completely.
//
// Table format ("sorted pack", built by the host once per code, models/layers.py:_Packed.sorted):
//   idx  [K][E] uint16  column t lists the neighbours of edge perm[t], valid entries first in the
//                       caller's order, unused slots = 0 (never dereferenced for their value)
//   cnt  [E]    uint8   number of valid entries of column t
//   perm [E]    uint16  edge of column t; columns are ordered by descending cnt so that the 32 edges
//                       of a warp have (nearly) the same list length.  NULL = identity.
// A warp loads and visits only slots below the largest cnt of its 32 columns: at BG2, 59 % of the
// check table and 55 % of the variable table is padding.  Per edge the valid neighbours are visited
// in the caller's order, so results do not depend on the permutation.
//
// Shared-memory arrays are row-interleaved, a[e * kRows + q]: the kRows codewords of the CTA share
// every gather address, so one LDS.64 / LDS.128 fetches the neighbour's value for all of them.

constexpr int kNeuralMaxL = 4;       // residual depths kept in registers; deeper queues use the per-layer kernels

template <int R>
__device__ __forceinline__ void neural_ldv(const float* p, float (&v)[R]) {
    if constexpr (R == 1) {
        v[0] = p[0];
    } else if constexpr (R == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        static_assert(R == 4, "kRows is 1, 2 or 4");
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}
template <int R>
__device__ __forceinline__ void neural_stv(float* p, const float (&v)[R]) {
    if constexpr (R == 1) {
        p[0] = v[0];
    } else if constexpr (R == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// sum over the valid slots of column t of src[idx][q]; all 32 lanes of the warp must call it
template <int kRows, int KV>
__device__ __forceinline__ void neural_gather_sum(const float* __restrict__ src, const unsigned short* __restrict__ col,
                                                  int Kv_rt, int E, int cnt, float (&acc)[kRows]) {
    const int K = KV ? KV : Kv_rt;
#pragma unroll
    for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
    const int kmax = __reduce_max_sync(0xffffffffu, cnt);
    if constexpr (KV > 0 && kNeuralPreload) {
        // the whole column of indices first (one L2 latency per edge instead of one per group) ...
        unsigned n[KV];
#pragma unroll
        for (int k = 0; k < KV; ++k) n[k] = (k / kNeuralGroup) * kNeuralGroup < kmax ? col[k * E] : 0u;
        // ... then groups of kNeuralGroup gathers between warp-uniform exits
#pragma unroll
        for (int k0 = 0; k0 < KV; k0 += kNeuralGroup) {
            if (k0 >= kmax) break;
            float v[kNeuralGroup][kRows];
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j)
                if (k0 + j < KV) neural_ldv<kRows>(src + n[k0 + j < KV ? k0 + j : 0] * kRows, v[j]);
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j)
                if (k0 + j < KV && k0 + j < cnt) {
#pragma unroll
                    for (int q = 0; q < kRows; ++q) acc[q] += v[j][q];
                }
        }
    } else {
#pragma unroll
        for (int k0 = 0; k0 < K; k0 += kNeuralGroup) {
            if (k0 >= kmax) break;                          // warp-uniform
            unsigned n[kNeuralGroup];
            float v[kNeuralGroup][kRows];
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j) n[j] = k0 + j < K ? col[(k0 + j) * E] : 0u;
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j) neural_ldv<kRows>(src + n[j] * kRows, v[j]);
#pragma unroll
            for (int j = 0; j < kNeuralGroup; ++j)
                if (k0 + j < cnt) {
#pragma unroll
                    for (int q = 0; q < kRows; ++q) acc[q] += v[j][q];
                }
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
    const unsigned char* __restrict__ ccnt, const unsigned short* __restrict__ cperm,
    const unsigned short* __restrict__ vidx, int Kv_rt, const unsigned char* __restrict__ vcnt,
    const unsigned short* __restrict__ vperm, const float* __restrict__ w_ch, const float* __restrict__ w_res, int L,
    int iters, long long B, int E, const float* __restrict__ gt, float* __restrict__ soft,
    float* __restrict__ max_loss) {
    extern __shared__ __align__(16) float sm[];
    const int Kc = KC ? KC : Kc_rt;
    const int Lb = L > 0 ? L : 1;
    const int RE = kRows * E;
    float* c2v = sm;                                  // [E][kRows] check-to-variable messages
    float* lls = sm + RE;                             // [E][kRows] channel LLRs of the resident codewords
    float* ring = sm + 2 * RE;                        // [Lb][E][kRows] earlier variable outputs
    __shared__ float red[kRows][kNeuralThreads / 32];
    const int lane = threadIdx.x & 31;
    float wr[kNeuralMaxL];
#pragma unroll
    for (int i = 0; i < kNeuralMaxL; ++i) wr[i] = i < L ? w_res[i] : 0.0f;

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
                lls[e * kRows + q] = v;
                ring[e * kRows + q] = v;
            }
        }
        __syncthreads();
        int cur = 0, nq = 0;                          // newest ring slot, queue length
        for (int l = 0; l < iters; ++l) {
            // ---- CheckLayer on x = ring[cur] ----
            const float* x = ring + cur * RE;
            // loads that do not depend on the gathers are issued a round early (cnt) or at the top of the round
            // (perm, w_ch): otherwise every round is a chain of three or four serialised L2 round trips
            int cnt_next = (int)threadIdx.x < E ? (int)ccnt[threadIdx.x] : 0;
            for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
                const bool live = t < E;
                const int tt = live ? t : 0;
                const int cnt = cnt_next;
                cnt_next = t + kNeuralThreads < E ? (int)ccnt[t + kNeuralThreads] : 0;
                const int e = cperm ? (int)cperm[tt] : tt;
                const unsigned short* col = cidx + tt;
                float mn[kRows];
                unsigned negb[kRows];
                bool zero[kRows];
#pragma unroll
                for (int q = 0; q < kRows; ++q) { mn[q] = CUDART_INF_F; negb[q] = 0u; zero[q] = false; }
                const int kmax = __reduce_max_sync(0xffffffffu, cnt);
                if constexpr (KC > 0 && kNeuralPreload) {
                    unsigned n[KC];
#pragma unroll
                    for (int k = 0; k < KC; ++k) n[k] = (k / kNeuralGroup) * kNeuralGroup < kmax ? col[k * E] : 0u;
#pragma unroll
                    for (int k0 = 0; k0 < KC; k0 += kNeuralGroup) {
                        if (k0 >= kmax) break;           // warp-uniform
                        float v[kNeuralGroup][kRows];
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j)
                            if (k0 + j < KC) neural_ldv<kRows>(x + n[k0 + j < KC ? k0 + j : 0] * kRows, v[j]);
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j)
                            if (k0 + j < KC && k0 + j < cnt) {
#pragma unroll
                                for (int q = 0; q < kRows; ++q) neural_check_visit(v[j][q], negb[q], zero[q], mn[q]);
                            }
                    }
                } else {
#pragma unroll
                    for (int k0 = 0; k0 < Kc; k0 += kNeuralGroup) {
                        if (k0 >= kmax) break;           // warp-uniform
                        unsigned n[kNeuralGroup];
                        float v[kNeuralGroup][kRows];
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j) n[j] = k0 + j < Kc ? col[(k0 + j) * E] : 0u;
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j) neural_ldv<kRows>(x + n[j] * kRows, v[j]);
#pragma unroll
                        for (int j = 0; j < kNeuralGroup; ++j)
                            if (k0 + j < cnt) {
#pragma unroll
                                for (int q = 0; q < kRows; ++q) neural_check_visit(v[j][q], negb[q], zero[q], mn[q]);
                            }
                    }
                }
                if (live) {
                    float o[kRows];
#pragma unroll
                    for (int q = 0; q < kRows; ++q) {
                        // a padded slot is a zero input: sign factor +1, magnitude 1e10 (layers.py:48-57)
                        const float m = (cnt < Kc && 1e10f < mn[q]) ? 1e10f : mn[q];
                        const float sp = __uint_as_float((negb[q] & 0x80000000u) | (zero[q] ? 0u : 0x3f800000u));
                        o[q] = sp * m;
                    }
                    neural_stv<kRows>(c2v + e * kRows, o);
                }
            }
            __syncthreads();
            if (l == iters - 1) break;
            // ---- VariableLayer(0, c2v) + ResidualLayer (layers.cuh neural_variable_fwd_kernel) ----
            const int nxt = nq == 0 ? 0 : (cur + 1) % Lb;   // empty slot, or the oldest entry once the queue is full
            int slot_off[kNeuralMaxL];
#pragma unroll
            for (int i = 0; i < kNeuralMaxL; ++i) slot_off[i] = ((cur - i + 2 * Lb) % Lb) * RE;
            int vcnt_next = (int)threadIdx.x < E ? (int)vcnt[threadIdx.x] : 0;
            for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
                const bool live = t < E;
                const int tt = live ? t : 0;
                const int cnt = vcnt_next;
                vcnt_next = t + kNeuralThreads < E ? (int)vcnt[t + kNeuralThreads] : 0;
                const int e = vperm ? (int)vperm[tt] : tt;
                const float w = w_ch[e];
                float acc[kRows];
                neural_gather_sum<kRows, KV>(c2v, vidx + tt, Kv_rt, E, cnt, acc);
                if (live) {
                    float r[kRows], ll[kRows], pv[kRows];
                    neural_ldv<kRows>(lls + e * kRows, ll);
#pragma unroll
                    for (int q = 0; q < kRows; ++q) r[q] = __fadd_rn(__fmul_rn(ll[q], w), acc[q]);
#pragma unroll
                    for (int i = 0; i < kNeuralMaxL; ++i)
                        if (i < nq) {
                            neural_ldv<kRows>(ring + slot_off[i] + e * kRows, pv);
#pragma unroll
                            for (int q = 0; q < kRows; ++q) r[q] = __fadd_rn(r[q], __fmul_rn(wr[i], pv[q]));
                        }
                    neural_stv<kRows>(ring + nxt * RE + e * kRows, r);
                }
            }
            cur = nxt;
            nq = nq < L ? nq + 1 : nq;
            __syncthreads();
        }
        // ---- final = VariableLayer(c2v, c2v); OutputLayer(final, llr, gt).  The ring is free now: soft values
        //      are staged there so that the global stores (and the ground-truth loads) are coalesced. ----
        float* stage = ring;
        int cnt_next = (int)threadIdx.x < E ? (int)vcnt[threadIdx.x] : 0;
        for (int t = threadIdx.x; t - lane < E; t += kNeuralThreads) {
            const bool live = t < E;
            const int tt = live ? t : 0;
            const int cnt = cnt_next;
            cnt_next = t + kNeuralThreads < E ? (int)vcnt[t + kNeuralThreads] : 0;
            const int e = vperm ? (int)vperm[tt] : tt;
            float acc[kRows];
            neural_gather_sum<kRows, KV>(c2v, vidx + tt, Kv_rt, E, cnt, acc);
            if (live) {
                float own[kRows], ll[kRows], s[kRows];
                neural_ldv<kRows>(c2v + e * kRows, own);
                neural_ldv<kRows>(lls + e * kRows, ll);
#pragma unroll
                for (int q = 0; q < kRows; ++q) {
                    const float z = __fadd_rn(__fadd_rn(own[q], acc[q]), ll[q]);
                    s[q] = 1.0f / (1.0f + expf(-z));
                }
                neural_stv<kRows>(stage + e * kRows, s);
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
                    const float s = stage[e * kRows + q];
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
                    float v = threadIdx.x < kNeuralThreads / 32 ? red[q][threadIdx.x] : -CUDART_INF_F;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (threadIdx.x == 0 && q < nb) max_loss[b0 + q] = v;
                }
            }
        }
    }
}

// ---- the same machinery for ONE layer at a time (training path, and the stand-alone CheckLayer / VariableLayer) ----
// kRows = 4 codewords per CTA, row-interleaved in shared memory (one LDS.128 per neighbour), sorted-pack tables.
// Results go straight to global memory at edge perm[t]: 4-byte scattered stores, merged by the L2.
#ifndef LDPC_PACKED_THREADS
#define LDPC_PACKED_THREADS 512
#endif
#ifndef LDPC_PACKED_CTAS
#define LDPC_PACKED_CTAS 2
#endif
constexpr int kPackedThreads = LDPC_PACKED_THREADS;
constexpr int kPackedRows = 4;

// CheckLayer.forward (layers.py:14-66).  nstar (optional) [B,E] int32: the edge whose |x| was selected as the
// minimum, -1 if the minimum is the 1e10 stand-in of a zero / padded input (no gradient, layers.py:55-58).
template <int KT>      // compile-time table width (9 for the BG2 check table), 0 = run-time K
__global__ void __launch_bounds__(kPackedThreads, LDPC_PACKED_CTAS) sorted_check_fwd_kernel(
    const float* __restrict__ x, const unsigned short* __restrict__ idx, int K_rt, const unsigned char* __restrict__ cnts,
    const unsigned short* __restrict__ perm, long long B, int E, float* __restrict__ out, int* __restrict__ nstar) {
    constexpr int R = kPackedRows;
    const int K = KT ? KT : K_rt;
    extern __shared__ __align__(16) float sm[];           // [E][R]
    const int lane = threadIdx.x & 31;
    for (long long b0 = (long long)blockIdx.x * R; b0 < B; b0 += (long long)gridDim.x * R) {
        const int nb = (int)((B - b0) < R ? (B - b0) : R);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const long long row = b0 + (q < nb ? q : nb - 1);
            for (int e = threadIdx.x; e < E; e += kPackedThreads) sm[e * R + q] = x[row * E + e];
        }
        __syncthreads();
        int cnt_next = (int)threadIdx.x < E ? (int)cnts[threadIdx.x] : 0;
        for (int t = threadIdx.x; t - lane < E; t += kPackedThreads) {
            const bool live = t < E;
            const int tt = live ? t : 0;
            const int cnt = cnt_next;
            cnt_next = t + kPackedThreads < E ? (int)cnts[t + kPackedThreads] : 0;
            const int e = perm ? (int)perm[tt] : tt;
            const unsigned short* col = idx + tt;
            float mn[R];
            unsigned negb[R];
            bool zero[R];
            int ns[R];
#pragma unroll
            for (int q = 0; q < R; ++q) { mn[q] = CUDART_INF_F; negb[q] = 0u; zero[q] = false; ns[q] = -1; }
            const int kmax = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += kNeuralGroup) {
                if (k0 >= kmax) break;                   // warp-uniform
                unsigned n[kNeuralGroup];
                float v[kNeuralGroup][R];
#pragma unroll
                for (int j = 0; j < kNeuralGroup; ++j) n[j] = k0 + j < K ? col[(k0 + j) * E] : 0u;
#pragma unroll
                for (int j = 0; j < kNeuralGroup; ++j) neural_ldv<R>(sm + n[j] * R, v[j]);
#pragma unroll
                for (int j = 0; j < kNeuralGroup; ++j)
                    if (k0 + j < cnt) {
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const float before = mn[q];
                            neural_check_visit(v[j][q], negb[q], zero[q], mn[q]);
                            if (mn[q] < before) ns[q] = v[j][q] != 0.0f ? (int)n[j] : -1;
                        }
                    }
            }
            if (live) {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (q < nb) {
                        const bool pad_wins = cnt < K && 1e10f < mn[q];
                        const float m = pad_wins ? 1e10f : mn[q];
                        const float sp = __uint_as_float((negb[q] & 0x80000000u) | (zero[q] ? 0u : 0x3f800000u));
                        out[(b0 + q) * E + e] = sp * m;
                        if (nstar) nstar[(b0 + q) * E + e] = pad_wins ? -1 : ns[q];
                    }
            }
        }
    }
}

// VariableLayer.forward (layers.py:78-125): out = llr + sum (w_ch == nullptr), or the variable + residual update
// of LDPCNeuralDecoder: out = (w_ch*llr + sum) + sum_i w_res[i]*prev[i]  (same operation order as the two layers).
template <int KT>      // 22 for the BG2 variable table, 0 = run-time K
__global__ void __launch_bounds__(kPackedThreads, LDPC_PACKED_CTAS) sorted_variable_fwd_kernel(
    const float* __restrict__ llr, const float* __restrict__ c2v, const unsigned short* __restrict__ idx, int K,
    const unsigned char* __restrict__ cnts, const unsigned short* __restrict__ perm, const float* __restrict__ w_ch,
    const float* __restrict__ w_res, ResidualPtrs prev, int L, long long B, int E, float* __restrict__ out) {
    constexpr int R = kPackedRows;
    extern __shared__ __align__(16) float sm[];           // [E][R]
    const int lane = threadIdx.x & 31;
    for (long long b0 = (long long)blockIdx.x * R; b0 < B; b0 += (long long)gridDim.x * R) {
        const int nb = (int)((B - b0) < R ? (B - b0) : R);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const long long row = b0 + (q < nb ? q : nb - 1);
            for (int e = threadIdx.x; e < E; e += kPackedThreads) sm[e * R + q] = c2v[row * E + e];
        }
        __syncthreads();
        int cnt_next = (int)threadIdx.x < E ? (int)cnts[threadIdx.x] : 0;
        for (int t = threadIdx.x; t - lane < E; t += kPackedThreads) {
            const bool live = t < E;
            const int tt = live ? t : 0;
            const int cnt = cnt_next;
            cnt_next = t + kPackedThreads < E ? (int)cnts[t + kPackedThreads] : 0;
            const int e = perm ? (int)perm[tt] : tt;
            const float w = w_ch ? w_ch[e] : 1.0f;
            float ll[R];                                   // channel LLRs in flight during the gathers
#pragma unroll
            for (int q = 0; q < R; ++q) ll[q] = llr[(b0 + (q < nb ? q : nb - 1)) * E + e];
            float acc[R];
            neural_gather_sum<R, KT>(sm, idx + tt, K, E, cnt, acc);
            if (live) {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (q < nb) {
                        const long long g = (b0 + q) * E + e;
                        float r;
                        if (w_ch) {
                            r = __fadd_rn(__fmul_rn(ll[q], w), acc[q]);
#pragma unroll
                            for (int i = 0; i < kMaxResidual; ++i)
                                if (i < L) r = __fadd_rn(r, __fmul_rn(__ldg(w_res + i), prev.prev[i][g]));
                        } else {
                            r = __fadd_rn(ll[q], acc[q]);
                        }
                        out[g] = r;
                    }
            }
        }
    }
}

// Backward of CheckLayer from (out, nstar): d out[e] / d x[nstar] = sign_product * sign(x[nstar]) and
// sign_product = sign(out[e]) (|out| = the selected |x| > 0, or out = +-0 when a sign factor was 0).
__global__ void __launch_bounds__(256) check_layer_bwd_nstar_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ out,
                                                                    const int* __restrict__ nstar,
                                                                    const float* __restrict__ grad_out, long long B,
                                                                    long long E, float* __restrict__ grad_x) {
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int n = nstar[t];
        if (n < 0) continue;
        const long long b = t / E;
        const float o = out[t], v = x[b * E + n];
        const float sp = o > 0.0f ? 1.0f : (o < 0.0f ? -1.0f : 0.0f);
        const float sv = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
        atomicAdd(&grad_x[b * E + n], grad_out[t] * sp * sv);
    }
}

}  // namespace ldpc
