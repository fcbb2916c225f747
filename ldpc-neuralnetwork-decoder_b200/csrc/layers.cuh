// layers.cuh -- edge-space check / variable / residual / output layers.
//
// Replaces models/layers.py: CheckLayer.forward (:14-66), VariableLayer.forward (:78-125),
// ResidualLayer.forward (:143-168), OutputLayer.forward (:180-210) and their autograd.
// Tensors are [B,E] fp32 row-major (E = Tanner edges, variable-major numbering of
// utils/ldpc_utils.py:77-84); idx is the [E,K] int64 neighbour table with -1 padding.
// The reference materialises (E,B,K) gathers; here one CTA stages a tile of `kRowsPerCta`
// codeword rows in shared memory and every thread walks its edge's neighbour list once for
// all staged rows, so the index table is read once per tile instead of once per codeword.
#pragma once
#include <math_constants.h>

#include "common.cuh"

#ifndef LDPC_LAYER_THREADS
#define LDPC_LAYER_THREADS 1024
#endif

namespace ldpc {

constexpr int kLayerThreads = LDPC_LAYER_THREADS;

// rows of x staged per CTA: as many as fit 200 KB, at most 8
inline int layer_rows_per_cta(long long E) {
    long long r = (200 * 1024) / (E * (long long)sizeof(float));
    return r < 1 ? 0 : (r > 8 ? 8 : (int)r);
}

// Neighbour-table accessor of these kernels: the reference layout ([E,K] int64, -1 padded).  A warp's load of slot k
// touches 32 rows K*8 bytes apart = 32 L1 wavefronts, which is what bounds them (ncu launch list r1: variable layer
// 606 us per 4096 x 6304 pass).  The fast path is csrc/neural.cuh: tables packed once per code to k-major uint16
// columns sorted by list length (ldpc_*_layer_fwd_sorted); these kernels remain for tables that cannot be packed
// (E >= 65535) and as the plain int64 ABI.
struct IdxI64 {
    const long long* p;
    __device__ __forceinline__ long long operator()(long long e, int k, long long, int K) const { return p[e * K + k]; }
};

template <int kRows, bool kStage, class Idx = IdxI64>
__global__ void __launch_bounds__(kLayerThreads) check_layer_fwd_kernel(const float* __restrict__ x,
                                                                        const Idx idx, long long B,
                                                                        long long E, int K, float* __restrict__ out,
                                                                        int* __restrict__ argmin_out) {
    extern __shared__ float xs[];   // [kRows][E] when staged
    for (long long b0 = (long long)blockIdx.x * kRows; b0 < B; b0 += (long long)gridDim.x * kRows) {
        const int nb = (int)((B - b0) < kRows ? (B - b0) : kRows);
        if constexpr (kStage) {
            __syncthreads();
            for (long long t = threadIdx.x; t < (long long)nb * E; t += kLayerThreads) xs[t] = x[b0 * E + t];
            __syncthreads();
        }
        for (long long e = threadIdx.x; e < E; e += kLayerThreads) {
            float sp[kRows], mn[kRows];
            int am[kRows];
#pragma unroll
            for (int q = 0; q < kRows; ++q) { sp[q] = 1.0f; mn[q] = CUDART_INF_F; am[q] = -1; }
            for (int k = 0; k < K; ++k) {
                const long long n = idx(e, k, E, K);
#pragma unroll
                for (int q = 0; q < kRows; ++q)
                    if (q < nb) {
                        // padded slots behave as value 0 (layers.py:48): sign(+1e-10)=+1, |0| -> 1e10
                        const float v = n < 0 ? 0.0f : (kStage ? xs[(long long)q * E + n] : x[(b0 + q) * E + n]);
                        const float sh = __fadd_rn(v, 1e-10f);
                        sp[q] *= sh > 0.0f ? 1.0f : (sh < 0.0f ? -1.0f : 0.0f);   // torch.sign, layers.py:52
                        float a = fabsf(v);
                        const bool real = a != 0.0f;
                        a = real ? a : 1e10f;                                       // layers.py:55-57
                        if (a < mn[q]) { mn[q] = a; am[q] = real ? k : -1; }
                    }
            }
#pragma unroll
            for (int q = 0; q < kRows; ++q)
                if (q < nb) {
                    out[(b0 + q) * E + e] = sp[q] * mn[q];
                    if (argmin_out) argmin_out[(b0 + q) * E + e] = am[q];
                }
        }
    }
}

// grad_x[b, idx[e,k*]] += grad_out[b,e] * sign_product * sign(x_k*); sign() has zero gradient,
// the in-place 0 -> 1e10 replacement blocks the gradient of zero entries (layers.py:55-58).
__global__ void __launch_bounds__(kLayerThreads) check_layer_bwd_kernel(const float* __restrict__ x,
                                                                        const long long* __restrict__ idx,
                                                                        const int* __restrict__ argmin,
                                                                        const float* __restrict__ grad_out, long long B,
                                                                        long long E, int K, float* __restrict__ grad_x) {
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / E, e = t - b * E;
        const int ks = argmin[t];
        if (ks < 0) continue;
        float sp = 1.0f;
        for (int k = 0; k < K; ++k) {
            const long long n = idx[e * K + k];
            const float v = n < 0 ? 0.0f : x[b * E + n];
            const float sh = __fadd_rn(v, 1e-10f);
            sp *= sh > 0.0f ? 1.0f : (sh < 0.0f ? -1.0f : 0.0f);
        }
        const long long n = idx[e * K + ks];
        const float v = x[b * E + n];
        const float sv = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
        atomicAdd(&grad_x[b * E + n], grad_out[t] * sp * sv);
    }
}

template <int kRows, bool kStage, class Idx = IdxI64>
__global__ void __launch_bounds__(kLayerThreads) variable_layer_fwd_kernel(const float* __restrict__ llr,
                                                                           const float* __restrict__ c2v,
                                                                           const Idx idx,
                                                                           long long B, long long E, int K,
                                                                           float* __restrict__ out) {
    extern __shared__ float xs[];
    for (long long b0 = (long long)blockIdx.x * kRows; b0 < B; b0 += (long long)gridDim.x * kRows) {
        const int nb = (int)((B - b0) < kRows ? (B - b0) : kRows);
        if constexpr (kStage) {
            __syncthreads();
            for (long long t = threadIdx.x; t < (long long)nb * E; t += kLayerThreads) xs[t] = c2v[b0 * E + t];
            __syncthreads();
        }
        for (long long e = threadIdx.x; e < E; e += kLayerThreads) {
            float acc[kRows];
#pragma unroll
            for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
            for (int k = 0; k < K; ++k) {
                const long long n = idx(e, k, E, K);
                if (n < 0) continue;
#pragma unroll
                for (int q = 0; q < kRows; ++q)
                    if (q < nb) acc[q] += kStage ? xs[(long long)q * E + n] : c2v[(b0 + q) * E + n];
            }
#pragma unroll
            for (int q = 0; q < kRows; ++q)
                if (q < nb) out[(b0 + q) * E + e] = __fadd_rn(llr[(b0 + q) * E + e], acc[q]);   // layers.py:123
        }
    }
}

__global__ void __launch_bounds__(kLayerThreads) variable_layer_bwd_kernel(const long long* __restrict__ idx,
                                                                           const float* __restrict__ grad_out,
                                                                           long long B, long long E, int K,
                                                                           float* __restrict__ grad_c2v) {
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / E, e = t - b * E;
        const float g = grad_out[t];
        for (int k = 0; k < K; ++k) {
            const long long n = idx[e * K + k];
            if (n >= 0) atomicAdd(&grad_c2v[b * E + n], g);
        }
    }
}

constexpr int kMaxResidual = 8;
struct ResidualPtrs {
    const float* prev[kMaxResidual];
};

// out = llr*w_ch + c2v, then + w_res[i]*prev[i] in list order (layers.py:157-166)
__global__ void __launch_bounds__(kLayerThreads) residual_layer_fwd_kernel(const float* __restrict__ llr,
                                                                           const float* __restrict__ c2v,
                                                                           const float* __restrict__ w_ch,
                                                                           const float* __restrict__ w_res,
                                                                           ResidualPtrs prev, int L, long long B,
                                                                           long long E, float* __restrict__ out) {
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long e = t % E;
        float r = __fadd_rn(__fmul_rn(llr[t], w_ch[e]), c2v[t]);
        for (int i = 0; i < L; ++i) r = __fadd_rn(r, __fmul_rn(w_res[i], prev.prev[i][t]));
        out[t] = r;
    }
}

// Variable + residual update of the unrolled neural min-sum decoder in one pass
// (notebook cell 11 `variable_layer_update`; the composing models/decoder.py is missing from the
// reference): out = (w_ch*llr + sum_k c2v[idx[e,k]]) + sum_{i<L} w_res[i]*prev[i], same operation
// order as VariableLayer(0, c2v) followed by ResidualLayer, so the result is bit-identical to the
// two-kernel composition while the gathered sum never makes a round trip through HBM.
template <int kRows, bool kStage, class Idx = IdxI64>
__global__ void __launch_bounds__(kLayerThreads) neural_variable_fwd_kernel(
    const float* __restrict__ llr, const float* __restrict__ c2v, const Idx idx,
    const float* __restrict__ w_ch, const float* __restrict__ w_res, ResidualPtrs prev, int L, long long B,
    long long E, int K, float* __restrict__ out) {
    extern __shared__ float xs[];
    float wr[kMaxResidual];
#pragma unroll
    for (int i = 0; i < kMaxResidual; ++i) wr[i] = i < L ? w_res[i] : 0.0f;
    for (long long b0 = (long long)blockIdx.x * kRows; b0 < B; b0 += (long long)gridDim.x * kRows) {
        const int nb = (int)((B - b0) < kRows ? (B - b0) : kRows);
        if constexpr (kStage) {
            __syncthreads();
            for (long long t = threadIdx.x; t < (long long)nb * E; t += kLayerThreads) xs[t] = c2v[b0 * E + t];
            __syncthreads();
        }
        for (long long e = threadIdx.x; e < E; e += kLayerThreads) {
            float acc[kRows];
#pragma unroll
            for (int q = 0; q < kRows; ++q) acc[q] = 0.0f;
            for (int k = 0; k < K; ++k) {
                const long long n = idx(e, k, E, K);
                if (n < 0) continue;
#pragma unroll
                for (int q = 0; q < kRows; ++q)
                    if (q < nb) acc[q] += kStage ? xs[(long long)q * E + n] : c2v[(b0 + q) * E + n];
            }
            const float w = w_ch[e];
#pragma unroll
            for (int q = 0; q < kRows; ++q)
                if (q < nb) {
                    const long long t = (b0 + q) * E + e;
                    float r = __fadd_rn(__fmul_rn(llr[t], w), acc[q]);
#pragma unroll
                    for (int i = 0; i < kMaxResidual; ++i)
                        if (i < L) r = __fadd_rn(r, __fmul_rn(wr[i], prev.prev[i][t]));
                    out[t] = r;
                }
        }
    }
}

// soft = sigmoid(final + llr); per-row max of BCE(soft, gt) with torch's log clamp at -100.
// One warp per row.
__global__ void __launch_bounds__(kLayerThreads) output_layer_fwd_kernel(const float* __restrict__ final_llr,
                                                                         const float* __restrict__ llr,
                                                                         const float* __restrict__ gt, long long B,
                                                                         long long E, float* __restrict__ soft,
                                                                         float* __restrict__ max_loss,
                                                                         int* __restrict__ argmax) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp0; b < B; b += nwarps) {
        float best = -CUDART_INF_F;
        int besti = 0x7fffffff;
        for (long long e = lane; e < E; e += 32) {
            const float z = __fadd_rn(final_llr[b * E + e], llr[b * E + e]);
            const float s = 1.0f / (1.0f + expf(-z));
            soft[b * E + e] = s;
            if (gt) {
                const float y = gt[b * E + e];
                const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                const float loss = -(y * l1 + (1.0f - y) * l0);
                if (loss > best) { best = loss; besti = (int)e; }
            }
        }
        if (gt) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            if (lane == 0) {
                max_loss[b] = best;
                if (argmax) argmax[b] = besti;
            }
        }
    }
}

inline int layer_grid(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ldpc
