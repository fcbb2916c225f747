// gnn_tc_pipe.cuh -- warp-specialised, software-pipelined edge kernel of the message-centred GNN decoder
// (MessageGNNLayer.forward, models/message_gnn_decoder.py:111-124: the message MLP over all E messages).
//
// Why a second edge kernel: ncu on gnn_edge_tc_kernel (gnn_tc.cuh) showed the tensor pipe 20 % active and every
// phase of a tile (load, split, MMA, epilogue, MMA, epilogue, store) serialised behind a CTA-wide barrier.
// Here the phases of neighbouring tiles overlap, and the activation operand never touches shared memory:
//
//   * loader warps (8): all global reads.  x rows and variable-node rows are requested one tile ahead into
//     registers, check-node rows with cp.async straight into their staging tile; every access is a coalesced
//     256-byte row (16 threads x 16 B) and lands XOR-swizzled (stage_ptr) so the per-row readers are conflict-free.
//   * row warps (8 = 2 per TMEM lane quarter): thread <-> message row.  They split comb = x + emb into hi/lo TF32
//     and write it to TENSOR MEMORY with tcgen05.st (the A operand of tcgen05.mma may live in TMEM: lane = row,
//     column = k; validated in tools/probe/umma_probe.cu), read the accumulator with tcgen05.ld, add the node
//     terms, apply ReLU, write the hidden activations back to TMEM as the next A operand, and finally stage and
//     store y with coalesced rows.
//   * MMA warp (1 thread issues): GEMM1 D1 = comb . W1A^T (K = 64, N = 128), GEMM2 D2 = h . W2^T (K = 128, N = 64),
//     each as 3xTF32 (lo.hi + hi.lo + hi.hi), weights (hi/lo canonical images, 128 KB) resident in shared memory.
//
// Tensor Memory map (512 columns): [0,64) comb hi | [64,128) comb lo | [128,256) D1 (D2 reuses [128,192)) |
//                                  [256,384) h hi | [384,512) h lo.
// Shared memory: 128 KB weights + 3 staging tiles of 32 KB: Sx (comb), Sp0 (variable-node rows, then y), Sp1 (check-node rows).
//
// Schedule per CTA (t = tile index of this CTA):
//   row warps:  E1(t)  C(t+1)  E2(t)  ST(t)          MMA thread:  G2(t)  G1(t+1)
// so the tensor pipe runs G2(t) while the row warps convert tile t+1, and G1(t+1) while they store tile t.
// All hand-offs are mbarriers (one phase per tile); every wait is bounded and traps instead of hanging the GPU.
#pragma once
#include "gnn_tc.cuh"

namespace ldpc {

constexpr int kPipeRowWarps = 8, kPipeLoaderWarps = 8;
constexpr int kPipeMmaWarp = kPipeRowWarps;                                    // warp 8
constexpr int kPipeThreads = (kPipeRowWarps + 1 + kPipeLoaderWarps) * 32;      // 544
constexpr int kPipeRowThreads = kPipeRowWarps * 32, kPipeLoaderThreads = kPipeLoaderWarps * 32;
constexpr size_t kPipeStage = 128 * 64 * sizeof(float);                        // 32 KB
constexpr size_t kPipeSmem = (size_t)(2 * 128 * 64 + 2 * 64 * 128) * sizeof(float) + 3 * kPipeStage;   // 224 KB
constexpr uint32_t kTmAcHi = 0, kTmAcLo = 64, kTmD1 = 128, kTmD2 = 128, kTmHHi = 256, kTmHLo = 384;

enum PipeBar { kBarFullX = 0, kBarFreeX, kBarAReady, kBarD1Full, kBarFullP, kBarFreeP1, kBarHReady, kBarD2Full, kBarFreeP0,
               kBarD2Drained, kNumPipeBars };

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// A operand in Tensor Memory (lane = row, column = k), B operand in shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_gemm3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K,
                                               uint32_t sbo_b, uint32_t idesc) {
    uint32_t acc = 0u;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t kb = (uint32_t)ks * 256u;                 // two 16-byte k-chunks of B per MMA; 8 TMEM columns of A
        umma_tf32_ts(d_tmem, a_lo + ks * 8, umma_desc(b_hi + kb, 128, sbo_b), idesc, acc);
        umma_tf32_ts(d_tmem, a_hi + ks * 8, umma_desc(b_lo + kb, 128, sbo_b), idesc, 1u);
        umma_tf32_ts(d_tmem, a_hi + ks * 8, umma_desc(b_hi + kb, 128, sbo_b), idesc, 1u);
        acc = 1u;
    }
}
// hi = x with the 13 low mantissa bits cleared (a TF32 value), lo = x - hi (exact; the tensor core reads its top 19
// bits).  Truncation instead of cvt.rna costs 2 instructions per element instead of 8 (cvt.rna.tf32.f32 expands to
// FSETP + VIADD + SEL + LOP3 on sm_100) and keeps the 3xTF32 product error at ~2^-20 relative.
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
        lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
    }
}

template <bool kResidual>
__global__ void __launch_bounds__(kPipeThreads, 1) gnn_edge_pipe_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l, const float* __restrict__ tc_l,
    const int* __restrict__ edge_var, const int* __restrict__ edge_chk, const int* __restrict__ edge_type,
    const float* __restrict__ Pv, const float* __restrict__ Pc, long long B, int E, int N, int M, float* __restrict__ y,
    int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* W1Ahi = tc_smem;                               // [128 x 64]
    uint8_t* W1Alo = W1Ahi + 128 * 64 * 4;
    uint8_t* W2hi = W1Alo + 128 * 64 * 4;                   // [64 x 128]
    uint8_t* W2lo = W2hi + 64 * 128 * 4;
    uint8_t* Sx = W2lo + 64 * 128 * 4;
    uint8_t* Sp0 = Sx + kPipeStage;
    uint8_t* Sp1 = Sp0 + kPipeStage;
    __shared__ uint64_t bars[kNumPipeBars];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float b2s[kH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == kPipeMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        auto init = [&](int b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bars[b])), "r"(count) : "memory"); };
        init(kBarFullX, kPipeLoaderWarps); init(kBarFreeX, kPipeRowWarps); init(kBarAReady, kPipeRowWarps); init(kBarD1Full, 1);
        init(kBarFullP, kPipeLoaderWarps); init(kBarFreeP1, kPipeRowWarps); init(kBarHReady, kPipeRowWarps); init(kBarD2Full, 1);
        init(kBarFreeP0, kPipeRowWarps); init(kBarD2Drained, kPipeRowWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // weights: canonical hi/lo images, contiguous in global memory in the same order as in shared memory
        const float4* src = reinterpret_cast<const float4*>(tc_l + kTcW1A);
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < (2 * 128 * 64 + 2 * 64 * 128) / 4; t += kPipeThreads) dst[t] = src[t];
        if (tid < kH) b2s[tid] = packed_l[kPkB2 + tid];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const long long rows = B * E, tiles = (rows + 127) / 128;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    bool ok = true;
    auto wait = [&](int b, uint32_t parity) { if (ok) ok = mbar_wait(&bars[b], parity & 1u); };
    // one elected lane arrives for the warp once every lane has finished the work the barrier publishes
    auto warp_arrive = [&](int b) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[b]); };

    if (warp < kPipeRowWarps) {
        // ================= row warps: thread <-> message row of the tile =================
        const int rowi = tid & 127, part = tid >> 7;                         // part 0: variable-node half, part 1: check-node half
        const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
        const int cc = tid & 15, cr0 = tid >> 4;                             // cooperative mapping of the y store
        const float4 bias = *reinterpret_cast<const float4*>(b2s + cc * 4);  // b2 of the 4 output columns this thread stores
        auto convert = [&]() {                                               // C: Sx -> comb hi/lo in TMEM, my 32 columns
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                const int col = part * 32 + c0;
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *stage_ptr(Sx, rowi, (col >> 2) + q);
                    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
                }
                uint32_t hi[16], lo[16];
                split16(v, hi, lo);
                tmem_st16(tmem + lane_base + kTmAcHi + col, hi);
                tmem_st16(tmem + lane_base + kTmAcLo + col, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(kBarFreeX);
            warp_arrive(kBarAReady);
        };
        if (my_tiles > 0) { wait(kBarFullX, 0); if (ok) convert(); }
        for (long long k = 0; k < my_tiles && ok; ++k) {
            const long long row0 = (blockIdx.x + k * gridDim.x) * 128;
            // E1: h = relu(D1 + node term) -> hi/lo in TMEM (my 64 hidden columns)
            wait(kBarD1Full, (uint32_t)k); wait(kBarFullP, (uint32_t)k);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                uint8_t* Sp = part == 0 ? Sp0 : Sp1;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    const int col = part * 64 + c0;
                    float h[16];
                    tmem_ld16(tmem + lane_base + kTmD1 + col, h);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 pq = *stage_ptr(Sp, rowi, (c0 >> 2) + q);
                        h[q * 4] = fmaxf(h[q * 4] + pq.x, 0.f); h[q * 4 + 1] = fmaxf(h[q * 4 + 1] + pq.y, 0.f);
                        h[q * 4 + 2] = fmaxf(h[q * 4 + 2] + pq.z, 0.f); h[q * 4 + 3] = fmaxf(h[q * 4 + 3] + pq.w, 0.f);
                    }
                    uint32_t hi[16], lo[16];
                    split16(h, hi, lo);
                    tmem_st16(tmem + lane_base + kTmHHi + col, hi);
                    tmem_st16(tmem + lane_base + kTmHLo + col, lo);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(kBarHReady);
            warp_arrive(kBarFreeP1);
            // C(t+1) while the tensor pipe runs GEMM2(t)
            if (k + 1 < my_tiles) { wait(kBarFullX, (uint32_t)(k + 1)); if (!ok) break; convert(); }
            // E2: D2 -> Sp0 (b2 and the residual are added by the store) (the variable-node rows of this tile are consumed: GEMM2 needed every row warp's h)
            wait(kBarD2Full, (uint32_t)k);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                const int col = part * 32 + c0;
                float o[16];
                tmem_ld16(tmem + lane_base + kTmD2 + col, o);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *stage_ptr(Sp0, rowi, (col >> 2) + q) = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(kBarD2Drained);                                      // GEMM1(t+1) may overwrite the accumulator columns
            asm volatile("bar.sync 1, %0;" :: "n"(kPipeRowThreads) : "memory");
            // ST: coalesced rows (+ residual x)
            {
                constexpr int kIters = 128 * 16 / kPipeRowThreads;
                float4 xv[kIters];
                if constexpr (kResidual) {
#pragma unroll
                    for (int it = 0; it < kIters; ++it) {
                        const long long row = row0 + it * (kPipeRowThreads / 16) + cr0;
                        xv[it] = row < rows ? reinterpret_cast<const float4*>(x + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    const int rr = it * (kPipeRowThreads / 16) + cr0;
                    float4 r = *stage_ptr(Sp0, rr, cc);
                    r.x += bias.x; r.y += bias.y; r.z += bias.z; r.w += bias.w;
                    if constexpr (kResidual) { r.x += xv[it].x; r.y += xv[it].y; r.z += xv[it].z; r.w += xv[it].w; }
                    if (row0 + rr < rows) reinterpret_cast<float4*>(y + (size_t)(row0 + rr) * kH)[cc] = r;
                }
            }
            warp_arrive(kBarFreeP0);
        }
    } else if (warp == kPipeMmaWarp) {
        // ================= MMA issue (one thread) =================
        if (lane == 0) {
            constexpr uint32_t kIdesc128 = umma_idesc_tf32(128), kIdesc64 = umma_idesc_tf32(64);
            auto gemm1 = [&]() {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                umma_gemm3_ts(tmem + kTmD1, tmem + kTmAcHi, tmem + kTmAcLo, smem_u32(W1Ahi), smem_u32(W1Alo), 64, 2048, kIdesc128);
                umma_commit(&bars[kBarD1Full]);
            };
            if (my_tiles > 0) { wait(kBarAReady, 0); if (ok) gemm1(); }
            for (long long k = 0; k < my_tiles && ok; ++k) {
                wait(kBarHReady, (uint32_t)k);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                umma_gemm3_ts(tmem + kTmD2, tmem + kTmHHi, tmem + kTmHLo, smem_u32(W2hi), smem_u32(W2lo), 128, 4096, kIdesc64);
                umma_commit(&bars[kBarD2Full]);
                if (k + 1 < my_tiles) {
                    wait(kBarAReady, (uint32_t)(k + 1)); wait(kBarD2Drained, (uint32_t)k);
                    if (!ok) break;
                    gemm1();
                }
            }
        }
    } else {
        // ================= loader warps: every global read, one tile ahead =================
        // Mapping: 4 threads per message row, thread q of them owns the 16-byte chunks {4j + q : j = 0..3}, so one warp instruction reads 64 contiguous bytes of 8 rows (full sectors) and
        // each thread serves only 2 rows per tile (rows lrow and lrow + 64): 2 index look-ups instead of 8.
        // Registers X hold the x rows of tile k+1 and V the variable-node rows of tile k+1 while tile k is computed,
        // and the node indices are fetched one tile ahead as well, so no hand-off to the row warps ever waits for a
        // dependent global-memory round trip.
        const int ltid = tid - (kPipeMmaWarp + 1) * 32;
        const int q4 = ltid & 3, lrow = ltid >> 2;                            // lrow in [0, 64)
        const long long stride = (long long)gridDim.x * 128;                  // rows between two tiles of this CTA
        const long long stride_b = stride / E;
        const int stride_e = (int)(stride % E);
        struct Pos { long long b0; int e0; };                                 // (codeword, message) of a tile's first row
        auto advance = [&](Pos& p) { p.b0 += stride_b; p.e0 += stride_e; if (p.e0 >= E) { p.e0 -= E; ++p.b0; } };
        float4 X[2][4], V[2][4];
        int ty[2], vrow[2], crow[2], crow_next[2];                            // type, Pv row, Pc row (-1: past the end)
        auto load_x = [&](long long k, const Pos& p) {                        // x rows + all indices of tile k
            const long long row0 = (blockIdx.x + k * gridDim.x) * 128;
            int ev[2], ec[2];
            long long bb[2];
            bool valid[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {                                     // every request first ...
                const int rr = lrow + 64 * h;
                int e = p.e0 + rr;
                long long b = p.b0;
                while (e >= E) { e -= E; ++b; }
                bb[h] = b;
                valid[h] = row0 + rr < rows;
                const float4* src = reinterpret_cast<const float4*>(x + (size_t)(row0 + rr) * kH);
#pragma unroll
                for (int j = 0; j < 4; ++j) X[h][j] = valid[h] ? src[4 * j + q4] : make_float4(0.f, 0.f, 0.f, 0.f);
                ty[h] = valid[h] ? __ldg(edge_type + e) : 0;
                ev[h] = valid[h] ? __ldg(edge_var + e) : 0;
                ec[h] = valid[h] ? __ldg(edge_chk + e) : 0;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {                                     // ... then the arithmetic that waits for them
                vrow[h] = valid[h] ? (int)(bb[h] * N) + ev[h] : -1;
                crow_next[h] = valid[h] ? (int)(bb[h] * M) + ec[h] : -1;
            }
        };
        auto load_v = [&]() {                                                 // variable-node rows (indices from load_x)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4* src = reinterpret_cast<const float4*>(Pv + (size_t)(vrow[h] < 0 ? 0 : vrow[h]) * kH);
#pragma unroll
                for (int j = 0; j < 4; ++j) V[h][j] = vrow[h] >= 0 ? src[4 * j + q4] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto publish_x = [&]() {                                              // comb = x + emb[type] -> Sx
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4* em = reinterpret_cast<const float4*>(emb_l + (size_t)ty[h] * kH);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 m = __ldg(em + 4 * j + q4);
                    *stage_ptr(Sx, lrow + 64 * h, 4 * j + q4) = make_float4(X[h][j].x + m.x, X[h][j].y + m.y, X[h][j].z + m.z, X[h][j].w + m.w);
                }
            }
            warp_arrive(kBarFullX);
        };
        Pos pos{((long long)blockIdx.x * 128) / E, (int)(((long long)blockIdx.x * 128) % E)};
        if (my_tiles > 0) {
            load_x(0, pos); load_v(); publish_x();
            crow[0] = crow_next[0]; crow[1] = crow_next[1];
        }
        for (long long k = 0; k < my_tiles && ok; ++k) {
            const bool more = k + 1 < my_tiles;
            advance(pos);
            if (more) load_x(k + 1, pos);                                     // in flight during everything below
            // check-node rows of tile k: asynchronous copies straight into Sp1 once E1(k-1) has released it
            wait(kBarFreeP1, (uint32_t)k + 1u);
            if (!ok) break;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* src = Pc + (size_t)(crow[h] < 0 ? 0 : crow[h]) * kH;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(stage_ptr(Sp1, lrow + 64 * h, 4 * j + q4))),
                                 "l"(src + (4 * j + q4) * 4), "r"(crow[h] >= 0 ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // variable-node rows of tile k (in registers since the previous iteration) -> Sp0 once y(k-1) has been stored
            wait(kBarFreeP0, (uint32_t)k + 1u);
            if (!ok) break;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) *stage_ptr(Sp0, lrow + 64 * h, 4 * j + q4) = V[h][j];
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            warp_arrive(kBarFullP);
            if (more) {
                load_v();
                crow[0] = crow_next[0]; crow[1] = crow_next[1];
                wait(kBarFreeX, (uint32_t)k);                                 // C(k) has consumed Sx
                if (!ok) break;
                publish_x();
            }
        }
    }
    if (!ok) { atomicExch(status, 1); asm volatile("trap;"); }               // a hand-off never arrived: fail loudly
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();
    if (warp == kPipeMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

}  // namespace ldpc
