// gnn_tc.cuh -- tensor-core (tcgen05) forward kernels of the message-centred GNN decoder.
//
// The two shared MLPs of MessageGNNLayer.forward (models/message_gnn_decoder.py:111-124) are a real
// dense contraction (hidden 64: per 128-message tile [128x64].[64x128] then [128x128].[128x64]), so
// they run on the 5th-generation tensor cores: tcgen05.mma kind::tf32, operands in shared memory
// (K-major, no-swizzle canonical layout), fp32 accumulators in Tensor Memory, completion through an
// mbarrier (tcgen05.commit), epilogue with tcgen05.ld (thread i of warp w owns TMEM lane 32w+i = row).
//
// Precision: a single TF32 product misses the 1e-4 tolerance (measured 2.6e-4 relative on a 64-deep
// dot product, tools/probe/umma_probe.cu).  Every operand is therefore split
// x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and each product is issued as three MMAs
// lo.hi + hi.lo + hi.hi ("3xTF32"): measured 5.5e-7 relative, i.e. fp32-level, at one third of the
// TF32 rate (still ~5x the FP32 FMA peak).
//
// This file holds the weight packing, the MMA / TMEM helpers, the node kernel and the single-buffered edge kernel
// (LDPC_GNN_EDGE=serial; the default edge kernel is the warp-specialised pipeline in gnn_tc_pipe.cuh).  Per
// 128-message tile the single-buffered edge kernel runs, each step behind a CTA-wide barrier:
//   1. comb = x + emb[type] -> split -> A_hi/A_lo (shared, canonical layout)
//   2. GEMM1: D1[128x128] = comb . W1A^T                       (8 k-steps x 3 MMAs, N=128)
//   3. h = relu(D1[:, 0:64] + Pv[var]) -> split -> A region;  GEMM2a: D2[128x64]  = h . W2[:, 0:64]^T
//   4. h = relu(D1[:, 64:128] + Pc[chk]) -> split -> A region; GEMM2b: D2 += h . W2[:, 64:128]^T
//   5. y = D2 + b2 (+ x for layers > 0)
// with every global access staged as coalesced 256-byte rows through a swizzled tile (stage_ptr).
// Shared memory: W1A hi/lo 64 KB + W2 hi/lo 64 KB + A region hi/lo 64 KB + staging 32 KB = 224 KB, TMEM 256 columns.
// The node kernel (Pv / Pc) has the same shape with a single GEMM per 128-node tile.
#pragma once
#include <cuda_fp16.h>

#include "gnn.cuh"

namespace ldpc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical no-swizzle K-major layout of a [R x K] fp32 operand: 8-row x 16-byte core matrices,
// K-adjacent core matrices 128 B apart (LBO), 8-row groups (K/4)*128 B apart (SBO)
__host__ __device__ constexpr uint32_t canon_off(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K / 4) * 128 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);      // version 1, SWIZZLE_NONE
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// the same with A = B = F16 (format 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// 3xTF32 product over K (multiple of 8): D (+)= (Ahi+Alo) . (Bhi+Blo)^T without the lo.lo term
__device__ __forceinline__ void umma_gemm3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int K, uint32_t lbo_a, uint32_t sbo_a, uint32_t sbo_b, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t ka = (uint32_t)ks * 2u * lbo_a, kb = (uint32_t)ks * 256u;  // two 16-byte k-chunks per MMA
        umma_tf32(d_tmem, umma_desc(a_lo + ka, lbo_a, sbo_a), umma_desc(b_hi + kb, 128, sbo_b), idesc, acc);
        umma_tf32(d_tmem, umma_desc(a_hi + ka, lbo_a, sbo_a), umma_desc(b_lo + kb, 128, sbo_b), idesc, 1u);
        umma_tf32(d_tmem, umma_desc(a_hi + ka, lbo_a, sbo_a), umma_desc(b_hi + kb, 128, sbo_b), idesc, 1u);
        acc = 1u;
    }
}
// bounded wait (never hang the GPU): returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    for (long long spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
        if (spins > 50000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Activation tiles ([128 x 64], written by the CUDA cores) of the shared-memory-operand kernels (node kernel and the
// single-buffered edge kernel).  Measured on B200 at B = 2048 (forward, 5 layers): a padded layout (LBO 144, legal: any
// multiple of 16 B) and prefetch.global.L2 of the node rows were both slower than this plain LBO 128 layout.
constexpr uint32_t kALbo = 128, kASbo = 16 * kALbo, kATileBytes = 16 * kASbo;
// write 4 consecutive k elements (one 16-byte chunk c) of row r, split hi/lo, into an activation tile pair
__device__ __forceinline__ void put_chunk(uint8_t* hi, uint8_t* lo, int r, int c, int K, float a, float b, float cc, float d) {
    (void)K;
    // activations: hi by truncation (1 LOP3 instead of the 4-instruction cvt.rna expansion), lo = exact remainder
    const float4 h = make_float4(tf32_trunc(a), tf32_trunc(b), tf32_trunc(cc), tf32_trunc(d));
    const float4 l = make_float4(a - h.x, b - h.y, cc - h.z, d - h.w);
    const uint32_t off = (uint32_t)((r >> 3) * kASbo + c * kALbo + (r & 7) * 16);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

// ---- weights: packed fp32 (gnn_pack_kernel) -> hi/lo canonical images in global memory -------------
// per layer: W1A [128 x 64] hi, lo | W2 [64 x 128] hi, lo | W1BV [64 x 64] hi, lo | W1BC [64 x 64] hi, lo  (floats)
//            | W2T [128 x 64] hi, lo (W2 transposed: B operand of dH = G . W2) | W1AT [64 x 128] hi, lo (B operand of
//              dcomb = dH . W1A) -- the two backward images (gnn_bwd_tc.cuh)
constexpr int kTcW1A = 0, kTcW2 = kTcW1A + 2 * 128 * 64, kTcW1BV = kTcW2 + 2 * 64 * 128, kTcW1BC = kTcW1BV + 2 * 64 * 64,
              kTcW2T = kTcW1BC + 2 * 64 * 64, kTcW1AT = kTcW2T + 2 * 128 * 64, kTcW1BVT = kTcW1AT + 2 * 64 * 128,
              kTcW1BCT = kTcW1BVT + 2 * 64 * 64, kTcPerLayer = kTcW1BCT + 2 * 64 * 64;      // + W1B transposed (dm = dP . W1B)
// fp16 two-way split images of the edge kernel's weights (gnn_tc_pipe.cuh, kF16): w = hi + lo with hi = fp16(w), lo = fp16(w - hi);
// canonical K-major layout for 16-bit elements: core matrix = 8 rows x 8 elements (16 B).  Per layer (halves):
// W1A [128 x 64] hi, lo | W2 [64 x 128] hi, lo = 64 KB, half of the TF32 images.
__host__ __device__ constexpr uint32_t canon_off16(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K / 8) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
constexpr int kTc16W1A = 0, kTc16W2 = 2 * 128 * 64, kTc16PerLayer = kTc16W2 + 2 * 64 * 128;
__global__ void gnn_pack_tc_kernel(const float* __restrict__ packed, float* __restrict__ tc, __half* __restrict__ tc16) {
    const int l = blockIdx.y;
    const float* pk = packed + (size_t)l * kPackedPerLayer;
    float* o = tc + (size_t)l * kTcPerLayer;
    __half* o16 = tc16 + (size_t)l * kTc16PerLayer;
    auto put16 = [&](__half* base, int rows, int K, int r, int k, float w) {
        const __half h = __float2half_rn(w);
        base[canon_off16(r, k, K) / 2] = h;
        base[rows * K + canon_off16(r, k, K) / 2] = __float2half_rn(w - __half2float(h));
    };
    auto put = [&](float* base, int rows, int K, int r, int k, float w) {
        const float h = tf32_hi(w);
        base[canon_off(r, k, K) / 4] = h;
        base[rows * K + canon_off(r, k, K) / 4] = tf32_hi(w - h);
    };
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 128 * 64; t += gridDim.x * blockDim.x) {
        put(o + kTcW1A, 128, 64, t / 64, t % 64, pk[kPkW1A + t]);                   // W1A[n][k], n < 128
        put(o + kTcW2, 64, 128, t / 128, t % 128, pk[kPkW2 + t]);                    // W2[n][k], k < 128
        put16(o16 + kTc16W1A, 128, 64, t / 64, t % 64, pk[kPkW1A + t]);
        put16(o16 + kTc16W2, 64, 128, t / 128, t % 128, pk[kPkW2 + t]);
        put(o + kTcW2T, 128, 64, t % 128, t / 128, pk[kPkW2 + t]);                   // W2T[k][n] = W2[n][k]
        put(o + kTcW1AT, 64, 128, t % 64, t / 64, pk[kPkW1A + t]);                   // W1AT[k][n] = W1A[n][k]
        if (t < 64 * 64) {
            put(o + kTcW1BV, 64, 64, t / 64, t % 64, pk[kPkW1BV + t]);
            put(o + kTcW1BC, 64, 64, t / 64, t % 64, pk[kPkW1BC + t]);
            put(o + kTcW1BVT, 64, 64, t % 64, t / 64, pk[kPkW1BV + t]);
            put(o + kTcW1BCT, 64, 64, t % 64, t / 64, pk[kPkW1BC + t]);
        }
    }
}

#ifndef GNN_TC_PARTS
#define GNN_TC_PARTS 4                 // edge kernel: threads per message row (each owns 1/PARTS of the columns)
#endif
constexpr int kEdgeParts = GNN_TC_PARTS, kEdgeThreads = 128 * kEdgeParts;
// 128 KB weight images + 64 KB activation tile (hi/lo) + 32 KB staging tile
constexpr size_t kEdgeTcSmem = (size_t)(2 * 128 * 64 + 2 * 64 * 128) * sizeof(float) + 2 * kATileBytes + 128 * 64 * sizeof(float);

// Staging tile S[128 rows][16 chunks of 16 B], chunk position XOR-swizzled by the row so that BOTH access
// patterns are bank-conflict free: the cooperative one (16 consecutive threads move the 16 chunks of one row =
// one coalesced 256-byte global access) and the per-row one (thread r reads/writes chunk c of its own row).
// 16-byte shared accesses are served per quarter warp (8 lanes -> the 8 bank groups of 16 B), so the swizzle must
// separate: 8 consecutive rows at one chunk (per-row readers), 8 consecutive chunks of one row (16-lane row movers)
// and 2 consecutive rows x 4 consecutive chunks (the 4-lane row movers of gnn_tc_pipe.cuh).  XOR with the row's low
// three bits rotated by one does all three.
__device__ __forceinline__ float4* stage_ptr(uint8_t* S, int r, int c) {
    return reinterpret_cast<float4*>(S + r * 256 + ((c ^ (((r & 1) << 2) | ((r >> 1) & 3))) << 4));
}

// kEdgeParts threads serve one message row for the per-row work (warp w, w+4, ... share TMEM lane quarter w%4 and
// split the columns).  Every global access of the kernel is a cooperative, fully coalesced 256-byte row transfer
// through S (ncu on the first version: L1 throughput was the busiest unit because per-row 16-byte accesses touch
// 32 different lines per request).
template <bool kResidual>
__global__ void __launch_bounds__(kEdgeThreads, 1) gnn_edge_tc_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l, const float* __restrict__ tc_l,
    const int* __restrict__ edge_var, const int* __restrict__ edge_chk, const int* __restrict__ edge_type,
    const float* __restrict__ Pv, const float* __restrict__ Pc, long long B, int E, int N, int M, float* __restrict__ y,
    int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* W1Ahi = tc_smem;                               // [128 x 64]
    uint8_t* W1Alo = W1Ahi + 128 * 64 * 4;
    uint8_t* W2hi = W1Alo + 128 * 64 * 4;                   // [64 x 128]
    uint8_t* W2lo = W2hi + 64 * 128 * 4;
    uint8_t* Ahi = W2lo + 64 * 128 * 4;                     // [128 x 64]: comb, then each half of relu(h)
    uint8_t* Alo = Ahi + kATileBytes;
    uint8_t* S = Alo + kATileBytes;                         // staging tile
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float b2s[kH];
    __shared__ int row_node[2][128];                        // variable / check node of each row of the tile
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rowi = tid & 127, part = tid >> 7;            // per-row work: row of the tile, column part
    constexpr int kColsPerPart = kH / kEdgeParts;
    constexpr int kCoopIters = 128 * 16 / kEdgeThreads;     // cooperative work: (row, chunk) pairs per thread
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // weights: the canonical hi/lo images are contiguous in global memory in the same order as in shared memory
        const float4* src = reinterpret_cast<const float4*>(tc_l + kTcW1A);
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < (2 * 128 * 64 + 2 * 64 * 128) / 4; t += kEdgeThreads) dst[t] = src[t];
        if (tid < kH) b2s[tid] = packed_l[kPkB2 + tid];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t d1 = tmem, d2 = tmem + 128;                                  // accumulators: columns [0,128) and [128,192)
    const uint32_t my_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    constexpr uint32_t kIdesc128 = umma_idesc_tf32(128), kIdesc64 = umma_idesc_tf32(64);
    uint32_t phase = 0;
    bool ok = true;
    const long long rows = B * E, tiles = (rows + 127) / 128;
    const int cc = tid & 15, cr0 = tid >> 4;                 // cooperative mapping: chunk, first row
    // software pipeline through registers: the x rows of the NEXT tile are requested while this tile computes
    float4 xr[kCoopIters];
    auto load_x = [&](long long t) {
#pragma unroll
        for (int it = 0; it < kCoopIters; ++it) {
            const long long row = t * 128 + it * (kEdgeThreads / 16) + cr0;
            xr[it] = (t < tiles && row < rows) ? reinterpret_cast<const float4*>(x + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_x(blockIdx.x);
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const long long row0 = tile * 128;
        const int e0 = (int)(row0 % E);                      // message index of the tile's first row
        const long long b0 = row0 / E;
        // 0. cooperative: comb = x + emb -> S;  node ids of the rows
        if (tid < 128) {
            int ee = e0 + tid; long long bb = b0;
            if (ee >= E) { ee -= E; bb += 1; }
            const bool lv = row0 + tid < rows;
            row_node[0][tid] = lv ? (int)(bb * N) + edge_var[ee] : -1;            // row of Pv (codeword-major)
            row_node[1][tid] = lv ? (int)(bb * M) + edge_chk[ee] : -1;
        }
#pragma unroll
        for (int it = 0; it < kCoopIters; ++it) {
            const int rr = it * (kEdgeThreads / 16) + cr0;
            int ee = e0 + rr; ee -= ee >= E ? E : 0;
            const float4 em = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[ee] * kH) + cc);
            *stage_ptr(S, rr, cc) = make_float4(xr[it].x + em.x, xr[it].y + em.y, xr[it].z + em.z, xr[it].w + em.w);
        }
        __syncthreads();
        // requests in flight behind the tensor-core work: both node-term row sets of this tile, x of the next tile
        float4 pn[2][kCoopIters];
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int it = 0; it < kCoopIters; ++it) {
                const int nd = row_node[half][it * (kEdgeThreads / 16) + cr0];
                pn[half][it] = nd >= 0 ? reinterpret_cast<const float4*>((half == 0 ? Pv : Pc) + (size_t)nd * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        load_x(tile + gridDim.x);
        // 1. per row: S -> A (split hi/lo)
#pragma unroll
        for (int c = 0; c < kColsPerPart / 4; ++c) {
            const int ch = part * (kColsPerPart / 4) + c;
            const float4 v = *stage_ptr(S, rowi, ch);
            put_chunk(Ahi, Alo, rowi, ch, 64, v.x, v.y, v.z, v.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // 2. GEMM1: D1 = comb . W1A^T;  meanwhile stage the variable-node terms of the tile in S (cooperative gather)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_gemm3(d1, smem_u32(Ahi), smem_u32(Alo), smem_u32(W1Ahi), smem_u32(W1Alo), 64, kALbo, kASbo, 2048, kIdesc128, false);
            umma_commit(&mbar);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (!ok) break;
#pragma unroll
            for (int it = 0; it < kCoopIters; ++it) *stage_ptr(S, it * (kEdgeThreads / 16) + cr0, cc) = pn[half][it];
            __syncthreads();
            ok = mbar_wait(&mbar, phase); phase ^= 1;        // GEMM1 (half 0) or GEMM2a (half 1) complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!ok) break;
#pragma unroll
            for (int c0 = 0; c0 < kColsPerPart; c0 += 16) {
                const int col = part * kColsPerPart + c0;
                float h[16];
                tmem_ld16(d1 + my_lane + half * kH + col, h);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 pq = *stage_ptr(S, rowi, (col >> 2) + q);
                    put_chunk(Ahi, Alo, rowi, (col >> 2) + q, 64, fmaxf(h[q * 4] + pq.x, 0.f), fmaxf(h[q * 4 + 1] + pq.y, 0.f),
                              fmaxf(h[q * 4 + 2] + pq.z, 0.f), fmaxf(h[q * 4 + 3] + pq.w, 0.f));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // W2 is [64 x 128] (SBO 4096); this half uses k in [64*half, 64*half+64): 16 chunks = 2048 bytes in
                umma_gemm3(d2, smem_u32(Ahi), smem_u32(Alo), smem_u32(W2hi) + half * 2048, smem_u32(W2lo) + half * 2048, 64,
                           kALbo, kASbo, 4096, kIdesc64, half != 0);
                umma_commit(&mbar);
            }
        }
        if (ok) { ok = mbar_wait(&mbar, phase); phase ^= 1; }   // GEMM2b complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // 5. per row: y = D2 + b2 -> S;  cooperative: (+ x) -> global
        if (ok) {
#pragma unroll
            for (int c0 = 0; c0 < kColsPerPart; c0 += 16) {
                const int col = part * kColsPerPart + c0;
                float o[16];
                tmem_ld16(d2 + my_lane + col, o);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *stage_ptr(S, rowi, (col >> 2) + q) = make_float4(o[q * 4] + b2s[col + q * 4], o[q * 4 + 1] + b2s[col + q * 4 + 1],
                                                                      o[q * 4 + 2] + b2s[col + q * 4 + 2], o[q * 4 + 3] + b2s[col + q * 4 + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                      // also: every thread is done reading TMEM before the next tile's MMAs
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
#pragma unroll
            for (int it = 0; it < kCoopIters; ++it) {
                const int rr = it * (kEdgeThreads / 16) + cr0;
                if (row0 + rr < rows) {
                    float4 r = *stage_ptr(S, rr, cc);
                    if constexpr (kResidual) {
                        const float4 v = reinterpret_cast<const float4*>(x + (size_t)(row0 + rr) * kH)[cc];
                        r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
                    }
                    reinterpret_cast<float4*>(y + (size_t)(row0 + rr) * kH)[cc] = r;
                }
            }
        }
        __syncthreads();                                      // S and row_node are rewritten by the next tile
    }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }     // MMA completion never arrived: fail loudly
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

}  // namespace ldpc
