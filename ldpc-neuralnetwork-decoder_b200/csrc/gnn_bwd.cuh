// gnn_bwd.cuh -- backward pass of the message-centred GNN decoder (training step).
//
// Replaces torch autograd through MessageGNNDecoder.forward (models/message_gnn_decoder.py:190-317)
// with the reference's training loss, ONE mean binary cross-entropy on the final probabilities
// (:313-315).  Gradients are produced for every parameter the reference's autograd reaches:
// input_embedding, every layer's type embeddings and both MLPs, and the LAST layer's
// output_projection; `output_layer` and the other layers' output_projection never receive a
// gradient in the reference and stay zero here.
//
// Forward (training=1) keeps the layer inputs x_0..x_L and the node terms Pv_l, Pc_l; hidden
// activations are recomputed.  Per layer, from the gradient G of the layer output:
//   edge kernel    hpre = W1A.comb + P[node]; dH = (W2^T.G) * [hpre>0]; dcomb = W1A^T.dH;
//                  dP[node] += dH (atomics); writes relu(hpre), dH, dcomb
//   outer kernel   dW2 += G^T.relu(h),  dW1A += dH^T.comb   (rows are the contraction dimension)
//   node kernel    dm = W1B^T.dP / deg;  writes m and dm;  outer kernel: dW1B += dP^T.m
//   finish kernel  dx = dcomb + dm_v[var] + dm_c[chk] (+ G for layers > 0); demb[type] += dx - [l>0]G
// fp32 FFMA throughout; weight gradients are accumulated in registers per CTA and flushed with
// one atomicAdd per entry.  This first version keeps relu(h)/dH in HBM between the kernels
// (6.4 MB per codeword-layer); fusing the outer products into the edge kernel is the next step.
#pragma once
#include "gnn.cuh"

namespace ldpc {

// packed gradient buffer of one layer: same layout as the packed weights (kPk*), plus
//   [kPackedPerLayer .. +types*h)  type-embedding gradient
// B2 slot receives d(b2v) = d(b2c); B1V/B1C the first-layer bias gradients.

// ---- loss + top gradient: one thread per (b, v) ------------------------------------------------
__global__ void __launch_bounds__(256) gnn_loss_kernel(const float* __restrict__ soft, const float* __restrict__ gt,
                                                        long long total, float* __restrict__ dsoft,
                                                        float* __restrict__ loss_out) {
    float part = 0.0f;
    const float inv = 1.0f / (float)total;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const float s = soft[t];
        const float p = 1.0f / (1.0f + expf(-s));
        const float y = gt[t];
        // torch.binary_cross_entropy: log terms clamped at -100; backward (p-y)/max(p(1-p),1e-12), then sigmoid'
        const float l1 = fmaxf(logf(p), -100.0f), l0 = fmaxf(logf(1.0f - p), -100.0f);
        part += -(y * l1 + (1.0f - y) * l0);
        const float pq = p * (1.0f - p);
        dsoft[t] = inv * (p - y) / fmaxf(pq, 1e-12f) * pq;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_out, part * inv);
}

// ---- readout backward: one thread per (b, e) ----------------------------------------------------
// G[b][e][:] = dsoft[b][var(e)] * w_out;  dw_out += ddec * x_L[e];  db_out += ddec
__global__ void __launch_bounds__(256) gnn_readout_bwd_kernel(const float* __restrict__ xL, const float* __restrict__ params,
                                                               GnnLayout lay, int last_layer, const float* __restrict__ dsoft,
                                                               const int* __restrict__ edge_var, long long B, int E, int N,
                                                               float* __restrict__ G, float* __restrict__ grad_params) {
    __shared__ __align__(16) float w[kH];
    __shared__ float acc[kH + 1];
    if (threadIdx.x < kH) w[threadIdx.x] = params[lay.out_w(last_layer) + threadIdx.x];
    if (threadIdx.x <= kH) acc[threadIdx.x] = 0.0f;
    __syncthreads();
    float gw[kH], gb = 0.0f;
#pragma unroll
    for (int k = 0; k < kH; ++k) gw[k] = 0.0f;
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t % E);
        const long long b = t / E;
        const float d = dsoft[b * N + edge_var[e]];
        const float4* xr = reinterpret_cast<const float4*>(xL + (size_t)t * kH);
        float4* gr = reinterpret_cast<float4*>(G + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) {
            const float4 a = xr[k4];
            gw[k4 * 4] = __fmaf_rn(d, a.x, gw[k4 * 4]); gw[k4 * 4 + 1] = __fmaf_rn(d, a.y, gw[k4 * 4 + 1]);
            gw[k4 * 4 + 2] = __fmaf_rn(d, a.z, gw[k4 * 4 + 2]); gw[k4 * 4 + 3] = __fmaf_rn(d, a.w, gw[k4 * 4 + 3]);
            gr[k4] = make_float4(d * w[k4 * 4], d * w[k4 * 4 + 1], d * w[k4 * 4 + 2], d * w[k4 * 4 + 3]);
        }
        gb += d;
    }
#pragma unroll
    for (int k = 0; k < kH; ++k) {
        float v = gw[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&acc[k], v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[kH], gb);
    __syncthreads();
    if (threadIdx.x < kH) atomicAdd(&grad_params[lay.out_w(last_layer) + threadIdx.x], acc[threadIdx.x]);
    if (threadIdx.x == kH) atomicAdd(&grad_params[lay.out_b(last_layer)], acc[kH]);
}

// out[n] += sum_k W[k*ldw + n] * a[k]  (transposed weights: W row-major [K][ldw]), k in [0,K)
template <int K, int NOUT>
__device__ __forceinline__ void matvecT(const float* __restrict__ W, int ldw, int n0, const float (&a)[K], float (&acc)[NOUT]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int n = 0; n < NOUT; n += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * ldw + n0 + n);
            acc[n] = __fmaf_rn(w.x, a[k], acc[n]); acc[n + 1] = __fmaf_rn(w.y, a[k], acc[n + 1]);
            acc[n + 2] = __fmaf_rn(w.z, a[k], acc[n + 2]); acc[n + 3] = __fmaf_rn(w.w, a[k], acc[n + 3]);
        }
    }
}

// ---- edge backward, part 1: one thread per (b, e) ------------------------------------------------
// recompute hpre = W1A.comb + P[node]; dH = (W2^T.G) * [hpre>0]; write relu(hpre), dH; dP[node] += dH.
// shared: W1A [2h][h], W2 [h][2h], and the thread's output-gradient row (gs, k-major, conflict-free).
__global__ void __launch_bounds__(kGnnThreads) gnn_edge_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l,
    const int* __restrict__ edge_var, const int* __restrict__ edge_chk, const int* __restrict__ edge_type,
    const float* __restrict__ Pv, const float* __restrict__ Pc, const float* __restrict__ G, long long B, int E, int N, int M,
    float* __restrict__ Hrelu, float* __restrict__ dH, float* __restrict__ dPv, float* __restrict__ dPc) {
    extern __shared__ __align__(16) float sm[];
    float* W1A = sm;                 // [2h][h]
    float* W2 = sm + 2 * kH * kH;    // [h][2h]
    float* gs = W2 + kH * 2 * kH + threadIdx.x;
    for (int t = threadIdx.x; t < 2 * kH * kH + kH * 2 * kH; t += kGnnThreads) sm[t] = packed_l[t];
    __syncthreads();
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * kGnnThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kGnnThreads) {
        const int e = (int)(t % E);
        const long long b = t / E;
        float a[kH];
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)t * kH);
        const float4* er = reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[e] * kH);
        const float4* gr = reinterpret_cast<const float4*>(G + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) {
            const float4 v = xr[k4], c = __ldg(er + k4), q = gr[k4];
            a[k4 * 4] = v.x + c.x; a[k4 * 4 + 1] = v.y + c.y; a[k4 * 4 + 2] = v.z + c.z; a[k4 * 4 + 3] = v.w + c.w;
            gs[(k4 * 4) * kGnnThreads] = q.x; gs[(k4 * 4 + 1) * kGnnThreads] = q.y;
            gs[(k4 * 4 + 2) * kGnnThreads] = q.z; gs[(k4 * 4 + 3) * kGnnThreads] = q.w;
        }
        const int var = edge_var[e], chk = edge_chk[e];
        const float4* pv = reinterpret_cast<const float4*>(Pv + ((size_t)b * N + var) * kH);
        const float4* pc = reinterpret_cast<const float4*>(Pc + ((size_t)b * M + chk) * kH);
        float* dpv = dPv + ((size_t)b * N + var) * kH;
        float* dpc = dPc + ((size_t)b * M + chk) * kH;
        float4* hr = reinterpret_cast<float4*>(Hrelu + (size_t)t * 2 * kH);
        float4* dhr = reinterpret_cast<float4*>(dH + (size_t)t * 2 * kH);
#pragma unroll 1
        for (int c0 = 0; c0 < 2 * kH; c0 += 16) {
            float h[16], dh[16];
            const float4* pp = c0 < kH ? pv + c0 / 4 : pc + (c0 - kH) / 4;
#pragma unroll
            for (int n = 0; n < 16; n += 4) {
                const float4 q = pp[n / 4];
                h[n] = q.x; h[n + 1] = q.y; h[n + 2] = q.z; h[n + 3] = q.w;
            }
            matvec64<16>(W1A + c0 * kH, kH, a, h);                       // pre-activations (as the forward)
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = 0.0f;
#pragma unroll 4
            for (int n = 0; n < kH; ++n) {                               // dh[j] = sum_n W2[n][c0+j] * g[n]
                const float gn = gs[n * kGnnThreads];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(W2 + n * 2 * kH + c0 + j);
                    dh[j] = __fmaf_rn(w.x, gn, dh[j]); dh[j + 1] = __fmaf_rn(w.y, gn, dh[j + 1]);
                    dh[j + 2] = __fmaf_rn(w.z, gn, dh[j + 2]); dh[j + 3] = __fmaf_rn(w.w, gn, dh[j + 3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const bool on = h[j] > 0.0f;
                h[j] = on ? h[j] : 0.0f;
                dh[j] = on ? dh[j] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                hr[(c0 + j) / 4] = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                dhr[(c0 + j) / 4] = make_float4(dh[j], dh[j + 1], dh[j + 2], dh[j + 3]);
            }
            float* dp = c0 < kH ? dpv + c0 : dpc + (c0 - kH);
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dp + j, dh[j]);
        }
    }
}

// ---- edge backward, part 2: dcomb[row][k] = sum_j W1A[j][k] * dH[row][j]; one thread per row ------
__global__ void __launch_bounds__(kGnnThreads) gnn_dcomb_kernel(const float* __restrict__ dH, const float* __restrict__ packed_l,
                                                                 long long rows, float* __restrict__ dcomb) {
    __shared__ __align__(16) float W1A[2 * kH * kH];
    for (int t = threadIdx.x; t < 2 * kH * kH; t += kGnnThreads) W1A[t] = packed_l[kPkW1A + t];
    __syncthreads();
    for (long long t = (long long)blockIdx.x * kGnnThreads + threadIdx.x; t < rows; t += (long long)gridDim.x * kGnnThreads) {
        float dc[kH];
#pragma unroll
        for (int k = 0; k < kH; ++k) dc[k] = 0.0f;
        const float4* dhr = reinterpret_cast<const float4*>(dH + (size_t)t * 2 * kH);
#pragma unroll 1
        for (int c0 = 0; c0 < 2 * kH; c0 += 16) {
            float dh[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 q = dhr[(c0 + j) / 4];
                dh[j] = q.x; dh[j + 1] = q.y; dh[j + 2] = q.z; dh[j + 3] = q.w;
            }
            matvecT<16, kH>(W1A + c0 * kH, kH, 0, dh, dc);
        }
        float4* dcr = reinterpret_cast<float4*>(dcomb + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) dcr[k4] = make_float4(dc[k4 * 4], dc[k4 * 4 + 1], dc[k4 * 4 + 2], dc[k4 * 4 + 3]);
    }
}

// ---- outer product accumulation: dW[NA][NB] += sum_r A[r][:]^T B[r][:] ----------------------------
// A rows [R][NA] (optionally A = x + emb[type] with kAddEmb), B rows [R][NB].  256 threads, each
// owns a (NA*NB/256)-entry tile in registers over the CTA's whole share of rows; also accumulates
// column sums of A (csumA, NA entries) if requested.  Flushed with atomics.
template <int NA, int NB, int kEmbOn>   // kEmbOn: 0 none, 2 = add emb[type(row)] to B rows
__global__ void __launch_bounds__(256) gnn_outer_kernel(const float* __restrict__ A, const float* __restrict__ Bm, long long R,
                                                         const float* __restrict__ emb_l, const int* __restrict__ edge_type,
                                                         int E, float* __restrict__ dW, float* __restrict__ csumA) {
    constexpr int TR = 32;                                  // rows per staged tile
    constexpr int TN = 4, TK = NA * NB / 256 / TN;          // per-thread tile TN (A cols) x TK (B cols)
    static_assert(NA * NB % (256 * TN) == 0 && TK % 4 == 0, "tile shape");
    __shared__ __align__(16) float As[TR][NA + 4];
    __shared__ __align__(16) float Bs[TR][NB + 4];
    const int tn = (threadIdx.x % (NA / TN)) * TN;          // A-column base
    const int tk = (threadIdx.x / (NA / TN)) * TK;          // B-column base
    float acc[TN][TK];
#pragma unroll
    for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = 0.0f;
    float cs = 0.0f;
    for (long long r0 = (long long)blockIdx.x * TR; r0 < R; r0 += (long long)gridDim.x * TR) {
        const int nr = (int)((R - r0) < TR ? (R - r0) : TR);
        __syncthreads();
        for (int t = threadIdx.x; t < TR * NA / 4; t += 256) {
            const int r = t / (NA / 4), c4 = t % (NA / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr) {
                v = reinterpret_cast<const float4*>(A + (size_t)(r0 + r) * NA)[c4];
            }
            *reinterpret_cast<float4*>(&As[r][c4 * 4]) = v;
        }
        for (int t = threadIdx.x; t < TR * NB / 4; t += 256) {
            const int r = t / (NB / 4), c4 = t % (NB / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr) {
                v = reinterpret_cast<const float4*>(Bm + (size_t)(r0 + r) * NB)[c4];
                if constexpr (kEmbOn == 2) {
                    const float4 c = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[(r0 + r) % E] * NB) + c4);
                    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
                }
            }
            *reinterpret_cast<float4*>(&Bs[r][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < TR; ++r) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[r][tn]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int j = 0; j < TK; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[r][tk + j]);
#pragma unroll
                for (int i = 0; i < TN; ++i) {
                    acc[i][j] = __fmaf_rn(av[i], b4.x, acc[i][j]); acc[i][j + 1] = __fmaf_rn(av[i], b4.y, acc[i][j + 1]);
                    acc[i][j + 2] = __fmaf_rn(av[i], b4.z, acc[i][j + 2]); acc[i][j + 3] = __fmaf_rn(av[i], b4.w, acc[i][j + 3]);
                }
            }
        }
        if (csumA && threadIdx.x < NA)
            for (int r = 0; r < nr; ++r) cs += As[r][threadIdx.x];
    }
#pragma unroll
    for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) atomicAdd(&dW[(size_t)(tn + i) * NB + tk + j], acc[i][j]);
    if (csumA && threadIdx.x < NA) atomicAdd(&csumA[threadIdx.x], cs);
}

// ---- node backward: one thread per (b, node) ----------------------------------------------------
// m = mean_{e in node}(x + emb) (recomputed, written to Mout); dm = W1B^T . dP / deg (written in place of dP? no:
// dP is still needed by the outer kernel, so dm goes to DMout).
__global__ void __launch_bounds__(kGnnThreads) gnn_node_bwd_kernel(const float* __restrict__ x, const float* __restrict__ emb_l,
                                                                    const float* __restrict__ packed_l, int kind,
                                                                    const int* __restrict__ ptr, const int* __restrict__ list,
                                                                    const int* __restrict__ edge_type, const float* __restrict__ dP,
                                                                    long long B, int E, int nodes, float* __restrict__ Mout,
                                                                    float* __restrict__ DMout) {
    __shared__ __align__(16) float Ws[kH * kH];            // W1B [n][k]
    const float* Wsrc = packed_l + (kind == 0 ? kPkW1BV : kPkW1BC);
    for (int t = threadIdx.x; t < kH * kH; t += kGnnThreads) Ws[t] = Wsrc[t];
    __syncthreads();
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
            const float4* er = reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[e] * kH);
#pragma unroll
            for (int k4 = 0; k4 < kH / 4; ++k4) {
                const float4 a = xr[k4], c = __ldg(er + k4);
                m[k4 * 4] += a.x + c.x; m[k4 * 4 + 1] += a.y + c.y; m[k4 * 4 + 2] += a.z + c.z; m[k4 * 4 + 3] += a.w + c.w;
            }
        }
        const float inv = 1.0f / (float)(k1 - k0);
        float4* mo = reinterpret_cast<float4*>(Mout + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4)
            mo[k4] = make_float4(m[k4 * 4] * inv, m[k4 * 4 + 1] * inv, m[k4 * 4 + 2] * inv, m[k4 * 4 + 3] * inv);
        // dm[k] = sum_n W1B[n][k] * dP[n]
        float dp[kH], dm[kH];
        const float4* dpr = reinterpret_cast<const float4*>(dP + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4) {
            const float4 q = dpr[k4];
            dp[k4 * 4] = q.x; dp[k4 * 4 + 1] = q.y; dp[k4 * 4 + 2] = q.z; dp[k4 * 4 + 3] = q.w;
            dm[k4 * 4] = 0.f; dm[k4 * 4 + 1] = 0.f; dm[k4 * 4 + 2] = 0.f; dm[k4 * 4 + 3] = 0.f;
        }
        matvecT<kH, kH>(Ws, kH, 0, dp, dm);
        float4* dmo = reinterpret_cast<float4*>(DMout + (size_t)t * kH);
#pragma unroll
        for (int k4 = 0; k4 < kH / 4; ++k4)
            dmo[k4] = make_float4(dm[k4 * 4] * inv, dm[k4 * 4 + 1] * inv, dm[k4 * 4 + 2] * inv, dm[k4 * 4 + 3] * inv);
    }
}

// ---- finish: dx = dcomb + dm_v[var] + dm_c[chk] (+ G if residual); demb[type] += dcomb_total ------
template <bool kResidual>
__global__ void __launch_bounds__(256) gnn_finish_bwd_kernel(const float* __restrict__ dcomb, const float* __restrict__ DMv,
                                                              const float* __restrict__ DMc, const float* __restrict__ G,
                                                              const int* __restrict__ edge_var, const int* __restrict__ edge_chk,
                                                              const int* __restrict__ edge_type, long long B, int E, int N, int M,
                                                              int types, float* __restrict__ dx, float* __restrict__ demb) {
    extern __shared__ float se[];                          // [types][h] block-local embedding gradient
    for (int t = threadIdx.x; t < types * kH; t += 256) se[t] = 0.0f;
    __syncthreads();
    const long long total = B * E * (kH / 4);
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
        const int q = (int)(t % (kH / 4));
        const long long be = t / (kH / 4);
        const int e = (int)(be % E);
        const long long b = be / E;
        float4 d = reinterpret_cast<const float4*>(dcomb)[t];
        const float4 mv = reinterpret_cast<const float4*>(DMv + ((size_t)b * N + edge_var[e]) * kH)[q];
        const float4 mc = reinterpret_cast<const float4*>(DMc + ((size_t)b * M + edge_chk[e]) * kH)[q];
        d.x += mv.x + mc.x; d.y += mv.y + mc.y; d.z += mv.z + mc.z; d.w += mv.w + mc.w;
        float* s = se + edge_type[e] * kH + q * 4;
        atomicAdd(s, d.x); atomicAdd(s + 1, d.y); atomicAdd(s + 2, d.z); atomicAdd(s + 3, d.w);
        if constexpr (kResidual) {
            const float4 g = reinterpret_cast<const float4*>(G)[t];
            d.x += g.x; d.y += g.y; d.z += g.z; d.w += g.w;
        }
        reinterpret_cast<float4*>(dx)[t] = d;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < types * kH; t += 256) atomicAdd(&demb[t], se[t]);
}

// ---- input embedding backward: dw_in[k] += sum dx0[e][k]*llr[var(e)], db_in[k] += sum dx0[e][k] ---
__global__ void __launch_bounds__(256) gnn_embed_bwd_kernel(const float* __restrict__ dx0, const float* __restrict__ llr,
                                                             const int* __restrict__ edge_var, long long B, int E, int N,
                                                             GnnLayout lay, float* __restrict__ grad_params) {
    __shared__ float sw[kH], sb[kH];
    if (threadIdx.x < kH) { sw[threadIdx.x] = 0.0f; sb[threadIdx.x] = 0.0f; }
    __syncthreads();
    // thread handles feature k = threadIdx.x % 64 of rows (blockIdx*4 + threadIdx/64), strided
    const int k = threadIdx.x % kH, lane_row = threadIdx.x / kH;
    float aw = 0.0f, ab = 0.0f;
    const long long total = B * E;
    for (long long t = (long long)blockIdx.x * 4 + lane_row; t < total; t += (long long)gridDim.x * 4) {
        const int e = (int)(t % E);
        const long long b = t / E;
        const float d = dx0[(size_t)t * kH + k];
        aw = __fmaf_rn(d, llr[b * N + edge_var[e]], aw);
        ab += d;
    }
    atomicAdd(&sw[k], aw);
    atomicAdd(&sb[k], ab);
    __syncthreads();
    if (threadIdx.x < kH) {
        atomicAdd(&grad_params[lay.in_w() + threadIdx.x], sw[threadIdx.x]);
        atomicAdd(&grad_params[lay.in_b() + threadIdx.x], sb[threadIdx.x]);
    }
}

// ---- packed gradients -> flat reference layout (+=) -----------------------------------------------
// pg: [kPackedPerLayer] packed-weight gradient + [types*h] embedding gradient, per layer
__global__ void gnn_unpack_grad_kernel(const float* __restrict__ pg, GnnLayout lay, int per_layer_pg,
                                       float* __restrict__ grad_params) {
    const int l = blockIdx.y;
    const float* g = pg + (size_t)l * per_layer_pg;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_layer_pg; t += gridDim.x * blockDim.x) {
        const float v = g[t];
        if (t < kPkW2) {
            const int n = t / kH, k = t % kH;
            if (n < kH) grad_params[lay.v_w1(l) + n * 2 * kH + k] += v;
            else grad_params[lay.c_w1(l) + (n - kH) * 2 * kH + k] += v;
        } else if (t < kPkB2) {
            const int u = t - kPkW2, n = u / (2 * kH), k = u % (2 * kH);
            if (k < kH) grad_params[lay.v_w2(l) + n * kH + k] += v;
            else grad_params[lay.c_w2(l) + n * kH + (k - kH)] += v;
        } else if (t < kPkW1BV) {
            const int n = t - kPkB2;                        // b2v and b2c receive the same gradient
            grad_params[lay.v_b2(l) + n] += v;
            grad_params[lay.c_b2(l) + n] += v;
        } else if (t < kPkW1BC) {
            const int u = t - kPkW1BV, n = u / kH, k = u % kH;
            grad_params[lay.v_w1(l) + n * 2 * kH + kH + k] += v;
        } else if (t < kPkB1V) {
            const int u = t - kPkW1BC, n = u / kH, k = u % kH;
            grad_params[lay.c_w1(l) + n * 2 * kH + kH + k] += v;
        } else if (t < kPkB1C) {
            grad_params[lay.v_b1(l) + (t - kPkB1V)] += v;
        } else if (t < kPackedPerLayer) {
            grad_params[lay.c_b1(l) + (t - kPkB1C)] += v;
        } else {
            grad_params[lay.emb(l) + (t - kPackedPerLayer)] += v;
        }
    }
}

}  // namespace ldpc
