// neural.cuh -- the unrolled neural min-sum decoder (LDPCNeuralDecoder) as ONE kernel.
//
// Replaces, for inference / validation, the per-iteration chain CheckLayer -> VariableLayer ->
// ResidualLayer and the final OutputLayer of models/layers.py (:14-66, :78-125, :143-168,
// :180-210) as composed by the reference's missing models/decoder.py (prototype:
// EE4002R_2025.ipynb cell 11 `forward`).  Every arithmetic step is the one of the per-layer
// kernels in layers.cuh, in the same order, so the output is bit-identical to the composition;
// what changes is where the state lives:
//   * a CTA keeps kRows codewords resident in shared memory for ALL iterations: the check
//     messages c2v[kRows][E] and a ring of max(L,1) earlier variable outputs x[kRows][E]
//     (x_l doubles as the input of the next check layer and as prev[0] of the next residual;
//     the new x overwrites the oldest ring slot in place -- each thread reads and writes only its
//     own element there).  Per codeword-iteration HBM sees one coalesced read of llr_e (25 KB at
//     BG2 Z=32) instead of six [B,E] round trips.
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
// the 5G BG2 tables (9 / 22) known, the k loops unroll completely and a thread has all its index
// loads in flight before the first shared-memory gather -- the kernel is latency-, not bandwidth-bound.
template <int kRows, int KC, int KV>
__global__ void __launch_bounds__(kNeuralThreads) neural_decode_kernel(
    const float* __restrict__ llr, const unsigned short* __restrict__ cidx, int Kc_rt,
    const unsigned short* __restrict__ vidx, int Kv_rt, const float* __restrict__ w_ch,
    const float* __restrict__ w_res, int L, int iters, long long B, int E, const float* __restrict__ gt,
    float* __restrict__ soft, float* __restrict__ max_loss) {
    extern __shared__ float sm[];
    const int Kc = KC ? KC : Kc_rt, Kv = KV ? KV : Kv_rt;
    const int Lb = L > 0 ? L : 1;
    float* c2v = sm;                                  // [kRows][E]
    float* ring = sm + (size_t)kRows * E;             // [Lb][kRows][E]
    __shared__ float red[kRows][kNeuralThreads / 32];

    for (long long b0 = (long long)blockIdx.x * kRows; b0 < B; b0 += (long long)gridDim.x * kRows) {
        const int nb = (int)((B - b0) < kRows ? (B - b0) : kRows);
        __syncthreads();
        // x_0 = llr_e sits in ring slot 0 but is NOT a queue entry (notebook cell 11: no residual in the first update)
        for (int q = 0; q < nb; ++q)
            for (int e = threadIdx.x; e < E; e += kNeuralThreads) ring[(size_t)q * E + e] = llr[(b0 + q) * E + e];
        __syncthreads();
        int cur = 0, nq = 0;                          // newest ring slot, queue length
        for (int l = 0; l < iters; ++l) {
            // ---- CheckLayer on x = ring[cur] (layers.cuh check_layer_fwd_kernel) ----
            const float* x = ring + (size_t)cur * kRows * E;
            for (int e = threadIdx.x; e < E; e += kNeuralThreads) {
                float sp[kRows], mn[kRows];
#pragma unroll
                for (int q = 0; q < kRows; ++q) { sp[q] = 1.0f; mn[q] = CUDART_INF_F; }
#pragma unroll
                for (int k = 0; k < Kc; ++k) {
                    const unsigned short n = cidx[(size_t)k * E + e];
#pragma unroll
                    for (int q = 0; q < kRows; ++q) {
                        const float v = n == kNeuralPad ? 0.0f : x[(size_t)q * E + n];
                        const float sh = __fadd_rn(v, 1e-10f);
                        sp[q] *= sh > 0.0f ? 1.0f : (sh < 0.0f ? -1.0f : 0.0f);
                        float a = fabsf(v);
                        a = a != 0.0f ? a : 1e10f;
                        mn[q] = a < mn[q] ? a : mn[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < kRows; ++q) c2v[(size_t)q * E + e] = sp[q] * mn[q];
            }
            __syncthreads();
            if (l == iters - 1) break;
            // ---- VariableLayer(0, c2v) + ResidualLayer (layers.cuh neural_variable_fwd_kernel) ----
            const int nxt = nq == 0 ? 0 : (cur + 1) % Lb;   // empty slot, or the oldest entry once the queue is full
            for (int e = threadIdx.x; e < E; e += kNeuralThreads) {
                float acc[kRows];
#pragma unroll
                for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
#pragma unroll
                for (int k = 0; k < Kv; ++k) {
                    const unsigned short n = vidx[(size_t)k * E + e];
                    if (n == kNeuralPad) continue;
#pragma unroll
                    for (int q = 0; q < kRows; ++q) acc[q] += c2v[(size_t)q * E + n];
                }
                const float w = w_ch[e];
#pragma unroll
                for (int q = 0; q < kRows; ++q)
                    if (q < nb) {
                        float r = __fadd_rn(__fmul_rn(llr[(b0 + q) * E + e], w), acc[q]);
                        for (int i = 0; i < nq; ++i) {
                            const int slot = (cur - i + Lb) % Lb;
                            r = __fadd_rn(r, __fmul_rn(__ldg(w_res + i), ring[((size_t)slot * kRows + q) * E + e]));
                        }
                        ring[((size_t)nxt * kRows + q) * E + e] = r;
                    }
            }
            cur = nxt;
            nq = nq < L ? nq + 1 : nq;
            __syncthreads();
        }
        // ---- final = VariableLayer(c2v, c2v); OutputLayer(final, llr, gt) ----
        float best[kRows];
#pragma unroll
        for (int q = 0; q < kRows; ++q) best[q] = -CUDART_INF_F;
        for (int e = threadIdx.x; e < E; e += kNeuralThreads) {
            float acc[kRows];
#pragma unroll
            for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
#pragma unroll
            for (int k = 0; k < Kv; ++k) {
                const unsigned short n = vidx[(size_t)k * E + e];
                if (n == kNeuralPad) continue;
#pragma unroll
                for (int q = 0; q < kRows; ++q) acc[q] += c2v[(size_t)q * E + n];
            }
#pragma unroll
            for (int q = 0; q < kRows; ++q)
                if (q < nb) {
                    const long long t = (b0 + q) * E + e;
                    const float fin = __fadd_rn(c2v[(size_t)q * E + e], acc[q]);
                    const float z = __fadd_rn(fin, llr[t]);
                    const float s = 1.0f / (1.0f + expf(-z));
                    soft[t] = s;
                    if (gt) {
                        const float y = gt[t];
                        const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                        const float loss = -(y * l1 + (1.0f - y) * l0);
                        best[q] = loss > best[q] ? loss : best[q];
                    }
                }
        }
        if (gt) {
#pragma unroll
            for (int q = 0; q < kRows; ++q) {
                float v = best[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
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
