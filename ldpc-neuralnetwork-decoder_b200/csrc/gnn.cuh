// gnn.cuh -- message-centred GNN decoder (models/message_gnn_decoder.py), inference path.
//
// Replaces MessageGNNLayer.forward (:51-129), MessageGNNDecoder.forward (:190-317) and the
// graph construction of TannerToMessageGraph (:356-536).  What the reference computes with two
// dense E x E normalised-adjacency bmm's per layer (:108,:118) is a per-variable-node and a
// per-check-node MEAN of the message features broadcast back to the node's messages (every row
// of D^-1/2 (A+I) D^-1/2 is 1/d on the d messages of the node), so the engine
// works on segment means and never builds an adjacency matrix.
//
// Per layer (features x[b][e][h], e = message in the reference's check-major / ascending
// variable order, h = hidden width 64):
//   comb = x + emb[type(e)]                                                  (:81-90)
//   node kernel   Pv[v] = W1v[:, h:] . mean_{e in v} comb + b1v               (:108-114)
//                 Pc[c] = W1c[:, h:] . mean_{e in c} comb + b1c               (:118-124)
//   edge kernel   hv = relu(W1v[:, :h] . comb + Pv[var(e)]),  hc likewise with Pc[chk(e)]
//                 out = W2v . hv + b2v + W2c . hc + b2c  (+ x for layers > 0)  (:127, :261-262)
// i.e. the first Linear of each MLP is split as W1.[comb ; m] = W1a.comb + W1b.m so the node
// half is evaluated once per node instead of once per message (fp32 reassociation only).
// Readout: dec[e] = w_out . x_L[e] + b_out, soft[v] = sum_{e in v} dec[e] (ascending check
// order, as the reference's message loop) + llr[v], prob = sigmoid(soft)    (:270-307).
//
// Arithmetic: fp32 FFMA.  One thread owns one message (or node) row with its 64 features in
// registers; weights are read from shared memory as warp-wide broadcasts (LDS.128 = 4 weights
// per instruction, 4 FFMA per load), hidden units are produced 16 at a time and consumed
// immediately by the second Linear, so no activation is staged in shared memory.  TF32/BF16
// tensor-core operands do not meet the 1e-4 tolerance; a 3xTF32
// tcgen05 variant is the planned replacement of the FFMA core.
#pragma once
#include <vector>

#include "common.cuh"
#include "tables.cuh"

struct ldpc_gnn {
    const ldpc_code* code = nullptr;
    int device = 0;
    int layers = 0, hidden = 0, types = 0;
    int E = 0, N = 0, M = 0;
    // device graph tables (message order = check-major, ascending variable)
    int* d_edge_var = nullptr;    // [E]
    int* d_edge_chk = nullptr;    // [E]
    int* d_edge_type = nullptr;   // [E]
    int* d_var_ptr = nullptr;     // [N+1]
    int* d_var_edge = nullptr;    // [E] messages of each variable, ascending check
    int* d_chk_ptr = nullptr;     // [M+1] (messages of a check are contiguous)
    int* d_var_list2 = nullptr;   // [E] message | type << 20 in variable-list order (gnn_node_pipe.cuh; null if E or types too large)
    int* d_chk_list2 = nullptr;   // [E] the same in check-list (= message) order
    float* d_packed = nullptr;    // per-call repacked weights (layers * kPackedPerLayer floats)
    float* d_emb = nullptr;       // per-call 16-byte aligned copy of the type embeddings [layers][types][h]
    void* d_tc16 = nullptr;       // per-call fp16 hi/lo images of the edge kernel's weights (gnn_tc.cuh, kTc16PerLayer halves per layer)
    float* d_tc = nullptr;        // per-call tf32 hi/lo weight images in the tensor-core operand layout (gnn_tc.cuh)
    int* d_status = nullptr;      // set to 1 if a tensor-core kernel timed out waiting for its MMAs
    size_t params = 0;
};

namespace ldpc {

constexpr int kH = 64;                       // hidden width the kernels are specialised for
constexpr int kGnnThreads = 128;
// flat parameter buffer, reference state_dict order
struct GnnLayout {
    int types;
    __host__ __device__ int in_w() const { return 0; }
    __host__ __device__ int in_b() const { return kH; }
    __host__ __device__ int per_layer() const { return types * kH + 2 * (kH * 2 * kH + kH + kH * kH + kH) + kH + 1; }
    __host__ __device__ int layer(int l) const { return 2 * kH + l * per_layer(); }
    __host__ __device__ int emb(int l) const { return layer(l); }
    __host__ __device__ int v_w1(int l) const { return emb(l) + types * kH; }
    __host__ __device__ int v_b1(int l) const { return v_w1(l) + kH * 2 * kH; }
    __host__ __device__ int v_w2(int l) const { return v_b1(l) + kH; }
    __host__ __device__ int v_b2(int l) const { return v_w2(l) + kH * kH; }
    __host__ __device__ int c_w1(int l) const { return v_b2(l) + kH; }
    __host__ __device__ int c_b1(int l) const { return c_w1(l) + kH * 2 * kH; }
    __host__ __device__ int c_w2(int l) const { return c_b1(l) + kH; }
    __host__ __device__ int c_b2(int l) const { return c_w2(l) + kH * kH; }
    __host__ __device__ int out_w(int l) const { return c_b2(l) + kH; }
    __host__ __device__ int out_b(int l) const { return out_w(l) + kH; }
    __host__ __device__ int total(int layers) const { return layer(layers) + kH + 1; }
};

// Repacked weights of one layer, laid out for broadcast LDS.128:
//   W1A [2h][h]   row n < h: W1v[n][0:h] (comb half), row n >= h: W1c[n-h][0:h]
//   W2  [h][2h]   row n: [W2v[n][0:h] | W2c[n][0:h]]
//   B2  [h]       b2v + b2c
//   W1BV [h][h], W1BC [h][h]   node halves W1v[n][h:2h], W1c[n][h:2h]
//   B1V [h], B1C [h]
constexpr int kPkW1A = 0, kPkW2 = kPkW1A + 2 * kH * kH, kPkB2 = kPkW2 + kH * 2 * kH, kPkW1BV = kPkB2 + kH,
              kPkW1BC = kPkW1BV + kH * kH, kPkB1V = kPkW1BC + kH * kH, kPkB1C = kPkB1V + kH,
              kPackedPerLayer = kPkB1C + kH;

__global__ void gnn_pack_kernel(const float* __restrict__ params, GnnLayout lay, int layers, float* __restrict__ packed,
                                float* __restrict__ emb_out) {
    const int l = blockIdx.y;
    float* o = packed + (size_t)l * kPackedPerLayer;
    // the flat parameter buffer gives odd offsets (26945 floats per layer): copy the embeddings to an aligned table
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < lay.types * kH; t += gridDim.x * blockDim.x)
        emb_out[(size_t)l * lay.types * kH + t] = params[lay.emb(l) + t];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < kPackedPerLayer; t += gridDim.x * blockDim.x) {
        float v;
        if (t < kPkW2) {
            const int n = t / kH, k = t % kH;
            v = n < kH ? params[lay.v_w1(l) + n * 2 * kH + k] : params[lay.c_w1(l) + (n - kH) * 2 * kH + k];
        } else if (t < kPkB2) {
            const int u = t - kPkW2, n = u / (2 * kH), k = u % (2 * kH);
            v = k < kH ? params[lay.v_w2(l) + n * kH + k] : params[lay.c_w2(l) + n * kH + (k - kH)];
        } else if (t < kPkW1BV) {
            const int n = t - kPkB2;
            v = params[lay.v_b2(l) + n] + params[lay.c_b2(l) + n];
        } else if (t < kPkW1BC) {
            const int u = t - kPkW1BV, n = u / kH, k = u % kH;
            v = params[lay.v_w1(l) + n * 2 * kH + kH + k];
        } else if (t < kPkB1V) {
            const int u = t - kPkW1BC, n = u / kH, k = u % kH;
            v = params[lay.c_w1(l) + n * 2 * kH + kH + k];
        } else if (t < kPkB1C) {
            v = params[lay.v_b1(l) + (t - kPkB1V)];
        } else {
            v = params[lay.c_b1(l) + (t - kPkB1C)];
        }
        o[t] = v;
    }
}

// x0[b][e][:] = w_in * llr[b][var(e)] + b_in     (message_gnn_decoder.py:218-235)
__global__ void __launch_bounds__(256) gnn_embed_kernel(const float* __restrict__ params, GnnLayout lay,
                                                         const float* __restrict__ llr, const int* __restrict__ edge_var,
                                                         long long B, int E, int N, float* __restrict__ x) {
    const long long total = B * E * (kH / 4);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t % (kH / 4));
        const long long be = t / (kH / 4);
        const int e = (int)(be % E);
        const long long b = be / E;
        const float v = llr[b * N + edge_var[e]];
        const float4 w = *reinterpret_cast<const float4*>(params + lay.in_w() + q * 4);
        const float4 bb = *reinterpret_cast<const float4*>(params + lay.in_b() + q * 4);
        float4 o;
        o.x = __fmaf_rn(v, w.x, bb.x); o.y = __fmaf_rn(v, w.y, bb.y); o.z = __fmaf_rn(v, w.z, bb.z); o.w = __fmaf_rn(v, w.w, bb.w);
        reinterpret_cast<float4*>(x)[be * (kH / 4) + q] = o;
    }
}

// acc[n] += sum_k W[n*ldw + k] * a[k], n in [0,NOUT), k in [0,64); W in shared memory, every lane
// reads the same address (broadcast LDS.128).
template <int NOUT>
__device__ __forceinline__ void matvec64(const float* __restrict__ W, int ldw, const float (&a)[kH], float (&acc)[NOUT]) {
#pragma unroll
    for (int n = 0; n < NOUT; ++n) {
        float s = acc[n];
#pragma unroll
        for (int k = 0; k < kH; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + n * ldw + k);
            s = __fmaf_rn(w.x, a[k], s); s = __fmaf_rn(w.y, a[k + 1], s);
            s = __fmaf_rn(w.z, a[k + 2], s); s = __fmaf_rn(w.w, a[k + 3], s);
        }
        acc[n] = s;
    }
}

// Node kernel: one thread per (codeword, node).  kind 0: variable nodes, 1: check nodes.
// P[b][node][:] = W1b . mean_{e in node}(x[e] + emb[type(e)]) + b1
__global__ void __launch_bounds__(kGnnThreads) gnn_node_kernel(const float* __restrict__ x, const float* __restrict__ emb_l,
                                                                const float* __restrict__ packed_l,
                                                                int kind, const int* __restrict__ ptr,
                                                                const int* __restrict__ list, const int* __restrict__ edge_type,
                                                                long long B, int E, int nodes, float* __restrict__ P) {
    __shared__ __align__(16) float Ws[kH * kH];
    __shared__ __align__(16) float bs[kH];
    const float* Wsrc = packed_l + (kind == 0 ? kPkW1BV : kPkW1BC);
    const float* bsrc = packed_l + (kind == 0 ? kPkB1V : kPkB1C);
    for (int t = threadIdx.x; t < kH * kH; t += kGnnThreads) Ws[t] = Wsrc[t];
    for (int t = threadIdx.x; t < kH; t += kGnnThreads) bs[t] = bsrc[t];
    __syncthreads();
    const float* emb = emb_l;
    const long long total = B * nodes;
    for (long long t = (long long)blockIdx.x * kGnnThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kGnnThreads) {
        const int node = (int)(t % nodes);
        const long long b = t / nodes;
        const int k0 = ptr[node], k1 = ptr[node + 1];
        float m[kH];
#pragma unroll
        for (int k = 0; k < kH; ++k) m[k] = 0.0f;
        for (int q = k0; q < k1; ++q) {
            const int e = list ? list[q] : q;
            const float4* xr = reinterpret_cast<const float4*>(x + ((size_t)b * E + e) * kH);
            const float4* er = reinterpret_cast<const float4*>(emb + (size_t)edge_type[e] * kH);
#pragma unroll
            for (int k4 = 0; k4 < kH / 4; ++k4) {
                const float4 a = xr[k4], c = __ldg(er + k4);
                m[k4 * 4] += a.x + c.x; m[k4 * 4 + 1] += a.y + c.y; m[k4 * 4 + 2] += a.z + c.z; m[k4 * 4 + 3] += a.w + c.w;
            }
        }
        const float inv = 1.0f / (float)(k1 - k0);
#pragma unroll
        for (int k = 0; k < kH; ++k) m[k] *= inv;
        float4* out = reinterpret_cast<float4*>(P + (size_t)t * kH);
#pragma unroll
        for (int c0 = 0; c0 < kH; c0 += 16) {
            float acc[16];
#pragma unroll
            for (int n = 0; n < 16; ++n) acc[n] = bs[c0 + n];
            matvec64<16>(Ws + c0 * kH, kH, m, acc);
#pragma unroll
            for (int n = 0; n < 16; n += 4) out[(c0 + n) / 4] = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
        }
    }
}

// Edge kernel: one thread per (codeword, message).
constexpr int kEdgeSmemFloats = 2 * kH * kH + kH * 2 * kH + kH;   // W1A + W2 + B2
template <bool kResidual>
__global__ void __launch_bounds__(kGnnThreads) gnn_edge_kernel(const float* __restrict__ x, const float* __restrict__ emb_l,
                                                                const float* __restrict__ packed_l,
                                                                const int* __restrict__ edge_var, const int* __restrict__ edge_chk,
                                                                const int* __restrict__ edge_type, const float* __restrict__ Pv,
                                                                const float* __restrict__ Pc, long long B, int E, int N, int M,
                                                                float* __restrict__ y) {
    extern __shared__ __align__(16) float sm[];
    float* W1A = sm;                       // [2h][h]
    float* W2 = sm + 2 * kH * kH;          // [h][2h]
    float* B2 = W2 + kH * 2 * kH;          // [h]
    for (int t = threadIdx.x; t < kEdgeSmemFloats; t += kGnnThreads) sm[t] = packed_l[t];
    __syncthreads();
    const float* emb = emb_l;
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * kGnnThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kGnnThreads) {
        const int e = (int)(t % E);
        const long long b = t / E;
        float a[kH], out[kH];
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)t * kH);
        const float4* er = reinterpret_cast<const float4*>(emb + (size_t)edge_type[e] * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) {
            const float4 v = xr[k4], c = __ldg(er + k4);
            a[k4 * 4] = v.x + c.x; a[k4 * 4 + 1] = v.y + c.y; a[k4 * 4 + 2] = v.z + c.z; a[k4 * 4 + 3] = v.w + c.w;
            if constexpr (kResidual) {
                out[k4 * 4] = v.x + B2[k4 * 4]; out[k4 * 4 + 1] = v.y + B2[k4 * 4 + 1];
                out[k4 * 4 + 2] = v.z + B2[k4 * 4 + 2]; out[k4 * 4 + 3] = v.w + B2[k4 * 4 + 3];
            } else {
                out[k4 * 4] = B2[k4 * 4]; out[k4 * 4 + 1] = B2[k4 * 4 + 1]; out[k4 * 4 + 2] = B2[k4 * 4 + 2]; out[k4 * 4 + 3] = B2[k4 * 4 + 3];
            }
        }
        const float4* pv = reinterpret_cast<const float4*>(Pv + ((size_t)b * N + edge_var[e]) * kH);
        const float4* pc = reinterpret_cast<const float4*>(Pc + ((size_t)b * M + edge_chk[e]) * kH);
        // hidden units 16 at a time: [0,64) variable-side MLP, [64,128) check-side MLP
#pragma unroll 1
        for (int c0 = 0; c0 < 2 * kH; c0 += 16) {
            float hcur[16];
            const float4* pp = c0 < kH ? pv + c0 / 4 : pc + (c0 - kH) / 4;
#pragma unroll
            for (int n = 0; n < 16; n += 4) {
                const float4 q = pp[n / 4];
                hcur[n] = q.x; hcur[n + 1] = q.y; hcur[n + 2] = q.z; hcur[n + 3] = q.w;
            }
            matvec64<16>(W1A + c0 * kH, kH, a, hcur);
#pragma unroll
            for (int n = 0; n < 16; ++n) hcur[n] = fmaxf(hcur[n], 0.0f);
            // out[n] += sum_j W2[n][c0 + j] * h[j]
#pragma unroll
            for (int n = 0; n < kH; ++n) {
                float s = out[n];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(W2 + n * 2 * kH + c0 + j);
                    s = __fmaf_rn(w.x, hcur[j], s); s = __fmaf_rn(w.y, hcur[j + 1], s);
                    s = __fmaf_rn(w.z, hcur[j + 2], s); s = __fmaf_rn(w.w, hcur[j + 3], s);
                }
                out[n] = s;
            }
        }
        float4* yr = reinterpret_cast<float4*>(y + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) yr[k4] = make_float4(out[k4 * 4], out[k4 * 4 + 1], out[k4 * 4 + 2], out[k4 * 4 + 3]);
    }
}

// Readout: one thread per (codeword, variable).
__global__ void __launch_bounds__(256) gnn_readout_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                           GnnLayout lay, int last_layer, const float* __restrict__ llr,
                                                           const int* __restrict__ var_ptr, const int* __restrict__ var_edge,
                                                           long long B, int E, int N, float* __restrict__ soft_out,
                                                           float* __restrict__ prob_out) {
    __shared__ __align__(16) float w[kH];
    __shared__ float bias;
    if (threadIdx.x < kH) w[threadIdx.x] = params[lay.out_w(last_layer) + threadIdx.x];
    if (threadIdx.x == 0) bias = params[lay.out_b(last_layer)];
    __syncthreads();
    const long long total = B * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t % N);
        const long long b = t / N;
        float s = 0.0f;
        for (int q = var_ptr[v]; q < var_ptr[v + 1]; ++q) {
            const float4* xr = reinterpret_cast<const float4*>(x + ((size_t)b * E + var_edge[q]) * kH);
            float d = 0.0f;
#pragma unroll
            for (int k4 = 0; k4 < kH / 4; ++k4) {
                const float4 a = xr[k4];
                d = __fmaf_rn(a.x, w[k4 * 4], d); d = __fmaf_rn(a.y, w[k4 * 4 + 1], d);
                d = __fmaf_rn(a.z, w[k4 * 4 + 2], d); d = __fmaf_rn(a.w, w[k4 * 4 + 3], d);
            }
            s += d + bias;                       // var_llrs[var] += decoded_llrs[msg]  (:289-298)
        }
        s += llr[t];                             // combined_llrs = var_llrs + input_llr  (:301)
        if (soft_out) soft_out[t] = s;
        if (prob_out) prob_out[t] = 1.0f / (1.0f + expf(-s));
    }
}

// Readout from per-message projections dec[b][e] = <x_L[e], w_out> (written by the last edge kernel, gnn_tc_pipe.cuh):
// soft[b][v] = llr + sum_{e in v} (dec[e] + b_out)   (message_gnn_decoder.py:289-301)
__global__ void __launch_bounds__(256) gnn_readout_sum_kernel(const float* __restrict__ dec, const float* __restrict__ params,
                                                               GnnLayout lay, int last_layer, const float* __restrict__ llr,
                                                               const int* __restrict__ var_ptr, const int* __restrict__ var_edge,
                                                               long long B, int E, int N, float* __restrict__ soft_out,
                                                               float* __restrict__ prob_out) {
    const float bias = params[lay.out_b(last_layer)];
    const long long total = B * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t % N);
        const long long b = t / N;
        float s = 0.0f;
        for (int q = var_ptr[v]; q < var_ptr[v + 1]; ++q) s += dec[(size_t)b * E + var_edge[q]] + bias;
        s += llr[t];
        if (soft_out) soft_out[t] = s;
        if (prob_out) prob_out[t] = 1.0f / (1.0f + expf(-s));
    }
}

inline int gnn_grid(long long items, int threads) {
    long long g = (items + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ldpc
