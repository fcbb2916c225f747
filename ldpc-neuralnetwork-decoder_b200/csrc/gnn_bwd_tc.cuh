// gnn_bwd_tc.cuh -- tensor-core (tcgen05, 3xTF32) data-gradient kernels of the GNN backward pass.
//
// The fp32 FFMA backward (gnn_bwd.cuh) spends 8.7 of its 15.6 ms per layer (B = 512, ncu launch list) in
// gnn_edge_bwd_kernel + gnn_dcomb_kernel, both at the FFMA ceiling (25-31 TFLOP/s).  Their three contractions are
// dense and move to the tensor cores here, with the same operand scheme as the forward (gnn_tc_pipe.cuh):
// activations split hi/lo by truncation and written to TENSOR MEMORY with tcgen05.st (A operand, lane = row),
// weights as hi/lo canonical K-major images in shared memory (B operand), fp32 accumulators in TMEM.
//
//   gnn_edge_bwd_tc_kernel   per 128-message tile:  D1 = comb . W1A^T (the forward's pre-activations, recomputed),
//                            DH = G . W2;  h = relu(D1 + P[node]) -> Hrelu;  dH = DH * [D1 + P > 0] -> dH, dP[node] += dH
//   gnn_dcomb_tc_kernel      dcomb = dH . W1A
// Autograd of MessageGNNLayer.forward (models/message_gnn_decoder.py:111-124).  The weight gradients (outer products
// over the rows) stay on the FFMA kernels: as MMAs they need MN-major operands, not yet validated in the probe.
//
// Every global access is a coalesced 256-byte row moved by 16 threads through the XOR-swizzled staging tiles
// (stage_ptr); dP uses 16-byte vector atomics (red.global.add.v4.f32).  All waits are bounded and trap.
#pragma once
#include "gnn_tc_pipe.cuh"

namespace ldpc {

constexpr int kBwdThreads = 512, kBwdParts = kBwdThreads / 128;
constexpr size_t kEdgeBwdTcSmem = (size_t)(2 * 128 * 64 + 2 * 128 * 64) * sizeof(float) + 2 * kPipeStage;     // 128 + 64 KB
constexpr size_t kDcombTcSmem = (size_t)(2 * 64 * 128) * sizeof(float) + 2 * kPipeStage;                       // 64 + 64 KB

__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1) gnn_edge_bwd_tc_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ tc_l,
    const int* __restrict__ edge_var, const int* __restrict__ edge_chk, const int* __restrict__ edge_type,
    const float* __restrict__ Pv, const float* __restrict__ Pc, const float* __restrict__ G, long long B, int E, int N, int M,
    float* __restrict__ Hrelu, float* __restrict__ dH, float* __restrict__ dPv, float* __restrict__ dPc, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* W1Ahi = tc_smem;                               // [128 x 64]
    uint8_t* W1Alo = W1Ahi + 128 * 64 * 4;
    uint8_t* W2Thi = W1Alo + 128 * 64 * 4;                  // [128 x 64] = W2 transposed
    uint8_t* W2Tlo = W2Thi + 128 * 64 * 4;
    uint8_t* S0 = W2Tlo + 128 * 64 * 4;                     // staging: comb, then Pv rows / h and dH columns [0,64)
    uint8_t* S1 = S0 + kPipeStage;                          // staging: G,    then Pc rows / h and dH columns [64,128)
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int row_node[2][128];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rowi = tid & 127, part = tid >> 7;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* a = reinterpret_cast<const float4*>(tc_l + kTcW1A);
        const float4* b = reinterpret_cast<const float4*>(tc_l + kTcW2T);
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < 2 * 128 * 64 / 4; t += kBwdThreads) { dst[t] = a[t]; dst[2 * 128 * 64 / 4 + t] = b[t]; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t my_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    // Tensor Memory: comb hi/lo [0,128) | D1 [128,256) | G hi/lo [256,384) | DH [384,512)
    constexpr uint32_t kAc = 0, kD1 = 128, kAg = 256, kDh = 384;
    constexpr uint32_t kIdesc128 = umma_idesc_tf32(128);
    uint32_t phase = 0;
    bool ok = true;
    const long long rows = B * E, tiles = (rows + 127) / 128;
    const int cc = tid & 15, cr0 = tid >> 4;                 // cooperative mapping: chunk, first row of the pass
    constexpr int kCoop = 128 * 16 / kBwdThreads, kStep = kBwdThreads / 16;
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const long long row0 = tile * 128;
        const int e0 = (int)(row0 % E);
        const long long b0 = row0 / E;
        if (tid < 128) {
            int ee = e0 + tid; long long bb = b0;
            while (ee >= E) { ee -= E; bb += 1; }
            const bool lv = row0 + tid < rows;
            row_node[0][tid] = lv ? (int)(bb * N) + edge_var[ee] : -1;
            row_node[1][tid] = lv ? (int)(bb * M) + edge_chk[ee] : -1;
        }
        // 0. cooperative: comb = x + emb -> S0, G -> S1
        {
            float4 xv[kCoop], gv[kCoop];
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const long long row = row0 + it * kStep + cr0;
                const bool lv = row < rows;
                xv[it] = lv ? reinterpret_cast<const float4*>(x + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
                gv[it] = lv ? reinterpret_cast<const float4*>(G + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const int rr = it * kStep + cr0;
                int ee = e0 + rr;
                while (ee >= E) ee -= E;
                const float4 em = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[ee] * kH) + cc);
                *stage_ptr(S0, rr, cc) = make_float4(xv[it].x + em.x, xv[it].y + em.y, xv[it].z + em.z, xv[it].w + em.w);
                *stage_ptr(S1, rr, cc) = gv[it];
            }
        }
        __syncthreads();
        // 1. per row: both operands -> Tensor Memory (my 16 columns of each)
        {
            float v[16];
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int op = 0; op < 2; ++op) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *stage_ptr(op == 0 ? S0 : S1, rowi, part * 4 + q);
                    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
                }
                split16(v, hi, lo);
                tmem_st16(tmem + my_lane + (op == 0 ? kAc : kAg) + part * 16, hi);
                tmem_st16(tmem + my_lane + (op == 0 ? kAc : kAg) + 64 + part * 16, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_gemm3_ts(tmem + kD1, tmem + kAc, tmem + kAc + 64, smem_u32(W1Ahi), smem_u32(W1Alo), 64, 2048, kIdesc128);
            umma_gemm3_ts(tmem + kDh, tmem + kAg, tmem + kAg + 64, smem_u32(W2Thi), smem_u32(W2Tlo), 64, 2048, kIdesc128);
            umma_commit(&mbar);
        }
        // meanwhile: node rows of the tile -> S0 (variable) / S1 (check)
        {
            float4 pv[kCoop], pc[kCoop];
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const int rr = it * kStep + cr0, nv = row_node[0][rr], nc = row_node[1][rr];
                pv[it] = nv >= 0 ? reinterpret_cast<const float4*>(Pv + (size_t)nv * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
                pc[it] = nc >= 0 ? reinterpret_cast<const float4*>(Pc + (size_t)nc * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                *stage_ptr(S0, it * kStep + cr0, cc) = pv[it];
                *stage_ptr(S1, it * kStep + cr0, cc) = pc[it];
            }
        }
        __syncthreads();
        ok = mbar_wait(&mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;
        // 2. per row, my 32 hidden columns: h = relu(D1 + P) -> staging in place; dH = DH * [h > 0] kept in registers
        uint8_t* Sp = part < 2 ? S0 : S1;
        const int pch = (part & 1) * 8;                      // first chunk of my columns in the staged row
        uint32_t da[16], db[16];
        {
            uint32_t ha[16], hb[16];
            tmem_ld16_issue(tmem + my_lane + kD1 + part * 32, ha);
            tmem_ld16_issue(tmem + my_lane + kD1 + part * 32 + 16, hb);
            tmem_ld_wait(ha, hb);
            tmem_ld16_issue(tmem + my_lane + kDh + part * 32, da);
            tmem_ld16_issue(tmem + my_lane + kDh + part * 32 + 16, db);
            tmem_ld_wait(da, db);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t (&hh)[16] = q < 4 ? ha : hb;
                uint32_t (&dd)[16] = q < 4 ? da : db;
                const float4 p = *stage_ptr(Sp, rowi, pch + q);
                const float add[4] = {p.x, p.y, p.z, p.w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float h = __uint_as_float(hh[(q & 3) * 4 + i]) + add[i];
                    const bool on = h > 0.0f;
                    o[i] = on ? h : 0.0f;
                    dd[(q & 3) * 4 + i] = on ? dd[(q & 3) * 4 + i] : 0u;
                }
                *stage_ptr(Sp, rowi, pch + q) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        // 3. cooperative: relu(h) rows -> Hrelu (512 B per row: columns [0,64) from S0, [64,128) from S1)
#pragma unroll
        for (int it = 0; it < kCoop; ++it) {
            const int rr = it * kStep + cr0;
            if (row0 + rr < rows) {
                float4* dst = reinterpret_cast<float4*>(Hrelu + (size_t)(row0 + rr) * 2 * kH);
                dst[cc] = *stage_ptr(S0, rr, cc);
                dst[16 + cc] = *stage_ptr(S1, rr, cc);
            }
        }
        __syncthreads();
        // 4. per row: dH -> staging;  cooperative: -> dH rows, dP[node] += dH (vector atomics)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t (&dd)[16] = q < 4 ? da : db;
            *stage_ptr(Sp, rowi, pch + q) = make_float4(__uint_as_float(dd[(q & 3) * 4]), __uint_as_float(dd[(q & 3) * 4 + 1]),
                                                        __uint_as_float(dd[(q & 3) * 4 + 2]), __uint_as_float(dd[(q & 3) * 4 + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int it = 0; it < kCoop; ++it) {
            const int rr = it * kStep + cr0;
            if (row0 + rr < rows) {
                const float4 a = *stage_ptr(S0, rr, cc), c = *stage_ptr(S1, rr, cc);
                float4* dst = reinterpret_cast<float4*>(dH + (size_t)(row0 + rr) * 2 * kH);
                dst[cc] = a;
                dst[16 + cc] = c;
                red_add_v4(dPv + (size_t)row_node[0][rr] * kH + cc * 4, a);
                red_add_v4(dPc + (size_t)row_node[1][rr] * kH + cc * 4, c);
            }
        }
        __syncthreads();
    }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// dcomb[row][:] = dH[row][:] . W1A   ([rows x 128] . [128 x 64]); one GEMM (K = 128, N = 64) per 128-row tile
__global__ void __launch_bounds__(kBwdThreads, 1) gnn_dcomb_tc_kernel(const float* __restrict__ dH, const float* __restrict__ tc_l,
                                                                     long long rows, float* __restrict__ dcomb,
                                                                     int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* Whi = tc_smem;                                 // W1AT [64 x 128]
    uint8_t* Wlo = Whi + 64 * 128 * 4;
    uint8_t* S0 = Wlo + 64 * 128 * 4;
    uint8_t* S1 = S0 + kPipeStage;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rowi = tid & 127, part = tid >> 7;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(tc_l + kTcW1AT);
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < 2 * 64 * 128 / 4; t += kBwdThreads) dst[t] = src[t];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t my_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    constexpr uint32_t kAhi = 0, kAlo = 128, kD = 256;       // Tensor Memory: dH hi [0,128) | dH lo [128,256) | D [256,320)
    constexpr uint32_t kIdesc64 = umma_idesc_tf32(64);
    uint32_t phase = 0;
    bool ok = true;
    const long long tiles = (rows + 127) / 128;
    const int cc = tid & 15, cr0 = tid >> 4;
    constexpr int kCoop = 128 * 16 / kBwdThreads, kStep = kBwdThreads / 16;
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const long long row0 = tile * 128;
        {
            float4 a[kCoop], c[kCoop];
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const long long row = row0 + it * kStep + cr0;
                const float4* src = reinterpret_cast<const float4*>(dH + (size_t)row * 2 * kH);
                a[it] = row < rows ? src[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
                c[it] = row < rows ? src[16 + cc] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                *stage_ptr(S0, it * kStep + cr0, cc) = a[it];
                *stage_ptr(S1, it * kStep + cr0, cc) = c[it];
            }
        }
        __syncthreads();
        {   // my 32 of the 128 k-columns: part 0,1 from S0, part 2,3 from S1
            uint8_t* Sp = part < 2 ? S0 : S1;
            const int pch = (part & 1) * 8;
            float v[16];
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *stage_ptr(Sp, rowi, pch + half * 4 + q);
                    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
                }
                split16(v, hi, lo);
                tmem_st16(tmem + my_lane + kAhi + part * 32 + half * 16, hi);
                tmem_st16(tmem + my_lane + kAlo + part * 32 + half * 16, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_gemm3_ts(tmem + kD, tmem + kAhi, tmem + kAlo, smem_u32(Whi), smem_u32(Wlo), 128, 4096, kIdesc64);
            umma_commit(&mbar);
        }
        ok = mbar_wait(&mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;
        {
            float o[16];
            tmem_ld16(tmem + my_lane + kD + part * 16, o);
#pragma unroll
            for (int q = 0; q < 4; ++q) *stage_ptr(S0, rowi, part * 4 + q) = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int it = 0; it < kCoop; ++it) {
            const int rr = it * kStep + cr0;
            if (row0 + rr < rows) reinterpret_cast<float4*>(dcomb + (size_t)(row0 + rr) * kH)[cc] = *stage_ptr(S0, rr, cc);
        }
        __syncthreads();
    }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ---- weight gradients on the tensor cores ---------------------------------------------------------------------
// dW[m][n] += sum_r A[r][m] * Bm[r][n]   (A rows 128 wide: relu(h) or dH;  Bm rows 64 wide: G or comb = x + emb)
// The contraction runs over the message rows, so both operands are MN-major for the MMA (K = row index).  The only
// shared-memory layout tcgen05.mma accepts for MN-major TF32 operands is SWIZZLE_128B_BASE32B (descriptor layout type
// 1; validated against fp64 in tools/probe/umma_probe.cu, `outer-mn mode=1`): atoms of 32 MN elements (128 B) x 4 K
// rows 128 B apart with the 32-byte chunk index XOR (row % 4), K groups 512 B apart (SBO), MN groups (R/4)*512 B
// apart (LBO).  One persistent CTA per SM keeps D[128 x 64] in Tensor Memory over all its tiles and flushes it with
// atomics once; the next tile's rows are requested into registers while the MMAs of the current tile run.
constexpr int kOuterThreads = 512;
constexpr uint32_t kOuterMnStride = (128 / 4) * 512;                                   // 16 KB per group of 32 MN elements
constexpr size_t kOuterTcSmem = (size_t)(2 * 128 * 128 + 2 * 128 * 64) * sizeof(float);   // A hi/lo 128 KB + B hi/lo 64 KB

__device__ __forceinline__ uint32_t outer_off(int q, int r) {                           // 16-byte chunk q (4 MN elements), row r
    return (uint32_t)((q >> 3) * kOuterMnStride + (r >> 2) * 512 + (r & 3) * 128 + ((((q & 7) >> 1) ^ (r & 3)) << 5) + (q & 1) * 16);
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((kOuterMnStride >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((512u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);   // version 1, SWIZZLE_128B_BASE32B
}

template <bool kAddEmb>
__global__ void __launch_bounds__(kOuterThreads, 1) gnn_outer_tc_kernel(
    const float* __restrict__ A, const float* __restrict__ Bm, long long rows, const float* __restrict__ emb_l,
    const int* __restrict__ edge_type, int E, float* __restrict__ dW, int stride_m, int stride_n, float* __restrict__ csumA,
    float* __restrict__ csumB, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* Ahi = tc_smem;
    uint8_t* Alo = Ahi + 128 * 128 * 4;
    uint8_t* Bhi = Alo + 128 * 128 * 4;
    uint8_t* Blo = Bhi + 128 * 64 * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // a_major = b_major = MN (bits 15, 16), D = F32, A = B = TF32, N = 64, M = 128
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // loader mapping: one warp instruction moves 4 rows x 128 contiguous bytes (8 chunks = one group of 32 MN elements);
    // lane = (chunk pair, row % 4, chunk parity), so that the 8 lanes of a quarter warp (2 chunks x 4 rows) hit the 8
    // distinct 16-byte slots ((chunk/2 ^ row%4) * 2 + chunk%2) of the swizzled 128-byte line: conflict-free stores
    // (ncu on the first mapping, 8 rows x 1 chunk per quarter warp: 2-way conflicts, L1 83 % busy)
    const int r3 = (lane >> 1) & 3, cq = (lane >> 3) * 2 + (lane & 1);
    const int qa = (warp & 3) * 8 + cq, ra0 = (warp >> 2) * 32 + r3;                    // A: 4 octets x 32 row quads, 8 per thread
    const int qb = (warp & 1) * 8 + cq, rb0 = (warp >> 1) * 16 + r3;                    // B: 2 octets x 32 row quads, 4 per thread
    const long long tiles = (rows + 127) / 128;
    float4 av[8], bv[4];
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    auto load_tile = [&](long long tile) {
        const long long row0 = tile * 128;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const long long row = row0 + ra0 + it * 4;
            av[it] = (tile < tiles && row < rows) ? reinterpret_cast<const float4*>(A + (size_t)row * 2 * kH)[qa] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const long long row = row0 + rb0 + it * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tile < tiles && row < rows) {
                v = reinterpret_cast<const float4*>(Bm + (size_t)row * kH)[qb];
                if constexpr (kAddEmb) {
                    const float4 em = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[(int)(row % E)] * kH) + qb);
                    v.x += em.x; v.y += em.y; v.z += em.z; v.w += em.w;
                }
            }
            bv[it] = v;
        }
    };
    uint32_t phase = 0;
    bool ok = true, first = true;
    load_tile(blockIdx.x);
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        if (!first) { ok = mbar_wait(&mbar, phase); phase ^= 1; if (!ok) break; }      // MMAs of the previous tile have read smem
        // registers -> hi/lo images (truncation split), column sums on the way
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const float4 v = av[it];
            const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
            const uint32_t o = outer_off(qa, ra0 + it * 4);
            *reinterpret_cast<float4*>(Ahi + o) = h;
            *reinterpret_cast<float4*>(Alo + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            sa.x += v.x; sa.y += v.y; sa.z += v.z; sa.w += v.w;
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const float4 v = bv[it];
            const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
            const uint32_t o = outer_off(qb, rb0 + it * 4);
            *reinterpret_cast<float4*>(Bhi + o) = h;
            *reinterpret_cast<float4*>(Blo + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            sb.x += v.x; sb.y += v.y; sb.z += v.z; sb.w += v.w;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t acc = first ? 0u : 1u;
            for (int ks = 0; ks < 16; ++ks) {                                         // 8 rows (two K groups of 4) per MMA
                const uint32_t ko = (uint32_t)ks * 1024u;
                const uint64_t ah = umma_desc_mn(smem_u32(Ahi) + ko), al = umma_desc_mn(smem_u32(Alo) + ko);
                const uint64_t bh = umma_desc_mn(smem_u32(Bhi) + ko), bl = umma_desc_mn(smem_u32(Blo) + ko);
                umma_tf32(tmem, al, bh, kIdesc, acc);
                umma_tf32(tmem, ah, bl, kIdesc, 1u);
                umma_tf32(tmem, ah, bh, kIdesc, 1u);
                acc = 1u;
            }
            umma_commit(&mbar);
        }
        first = false;
        load_tile(tile + gridDim.x);                                                  // in flight while the tensor pipe works
    }
    if (!first && ok) { ok = mbar_wait(&mbar, phase); phase ^= 1; }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!first && warp < 4) {                                                         // lane of TMEM = m
        const int m = tid;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
            float o[16];
            tmem_ld16(tmem + (((uint32_t)(warp * 32)) << 16) + c0, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dW + (size_t)m * stride_m + (size_t)(c0 + j) * stride_n, o[j]);
        }
    }
    // column sums: the 4 lanes with the same chunk (row % 4 = 0..3) first, then one atomic per chunk and warp
#pragma unroll
    for (int s = 2; s < 8; s <<= 1) {
        sa.x += __shfl_xor_sync(0xffffffffu, sa.x, s); sa.y += __shfl_xor_sync(0xffffffffu, sa.y, s);
        sa.z += __shfl_xor_sync(0xffffffffu, sa.z, s); sa.w += __shfl_xor_sync(0xffffffffu, sa.w, s);
        sb.x += __shfl_xor_sync(0xffffffffu, sb.x, s); sb.y += __shfl_xor_sync(0xffffffffu, sb.y, s);
        sb.z += __shfl_xor_sync(0xffffffffu, sb.z, s); sb.w += __shfl_xor_sync(0xffffffffu, sb.w, s);
    }
    if (r3 == 0) {
        if (csumA) red_add_v4(csumA + qa * 4, sa);
        if (csumB) red_add_v4(csumB + qb * 4, sb);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

// ---- node backward on the tensor cores: dm[row][:] = (dP[row][:] . W1B) / deg(node) -------------------------------
// (the node means themselves were saved by the training forward, gnn_node_tc_kernel's Msave, so no gather is repeated)
__global__ void __launch_bounds__(kNodeThreads, 2) gnn_node_dm_tc_kernel(const float* __restrict__ dP, const float* __restrict__ tc_l,
                                                                        int kind, const int* __restrict__ ptr, long long rows,
                                                                        int nodes, float* __restrict__ DMout, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* Whi = tc_smem;                                 // W1B transposed [64 x 64]
    uint8_t* Wlo = Whi + 64 * 64 * 4;
    uint8_t* S = Wlo + 64 * 64 * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rowi = tid & 127, part = tid >> 7;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(tc_l + (kind == 0 ? kTcW1BVT : kTcW1BCT));
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < 2 * 64 * 64 / 4; t += kNodeThreads) dst[t] = src[t];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t my_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    constexpr uint32_t kIdesc64 = umma_idesc_tf32(64);
    uint32_t phase = 0;
    bool ok = true;
    const long long tiles = (rows + 127) / 128;
    const int cc = tid & 15, cr0 = tid >> 4;
    constexpr int kCoop = 128 * 16 / kNodeThreads, kStep = kNodeThreads / 16;
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const long long row0 = tile * 128;
        {
            float4 v[kCoop];
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const long long row = row0 + it * kStep + cr0;
                v[it] = row < rows ? reinterpret_cast<const float4*>(dP + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < kCoop; ++it) *stage_ptr(S, it * kStep + cr0, cc) = v[it];
        }
        __syncthreads();
        {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = *stage_ptr(S, rowi, part * 4 + q);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
            uint32_t hi[16], lo[16];
            split16(v, hi, lo);
            tmem_st16(tmem + my_lane + kTmNodeAHi + part * 16, hi);
            tmem_st16(tmem + my_lane + kTmNodeALo + part * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_gemm3_ts(tmem + kTmNodeD, tmem + kTmNodeAHi, tmem + kTmNodeALo, smem_u32(Whi), smem_u32(Wlo), 64, 2048, kIdesc64);
            umma_commit(&mbar);
        }
        ok = mbar_wait(&mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const long long row = row0 + rowi;
            float inv = 0.0f;
            if (row < rows) { const int node = (int)(row % nodes); inv = 1.0f / (float)(ptr[node + 1] - ptr[node]); }
            float o[16];
            tmem_ld16(tmem + my_lane + kTmNodeD + part * 16, o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *stage_ptr(S, rowi, part * 4 + q) = make_float4(o[q * 4] * inv, o[q * 4 + 1] * inv, o[q * 4 + 2] * inv, o[q * 4 + 3] * inv);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
#pragma unroll
            for (int it = 0; it < kCoop; ++it) {
                const int rr = it * kStep + cr0;
                if (row0 + rr < rows) reinterpret_cast<float4*>(DMout + (size_t)(row0 + rr) * kH)[cc] = *stage_ptr(S, rr, cc);
            }
        }
        __syncthreads();
    }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

}  // namespace ldpc
