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
// (Measured alternative: h hi written in place over D1 with D2 in its own columns, so that GEMM1(t+1) can be queued
// right behind GEMM2(t): 2-3 ms SLOWER per forward at B = 2048 -- the accumulator drain, the in-place operand and the
// MMAs then compete for the same Tensor Memory columns.)
// Shared memory: 128 KB weights + 3 staging tiles of 32 KB: Sx (comb), Sp0 (variable-node rows, then y), Sp1 (check-node rows).
//
// Schedule per CTA (t = tile index of this CTA):
//   row warps:  E1(t)  C(t+1)  E2(t)  ST(t)          MMA thread:  G2(t) first half | second half   G1(t+1)
// so the tensor pipe runs G2(t) while the row warps convert tile t+1, and G1(t+1) while they store tile t.
// (Also measured: starting G2(t) on the first half of the hidden columns while E1(t) produces the rest, which needs
// the D2-in-own-columns map above: 35.4 ms.)
// All hand-offs are mbarriers (one phase per tile); every wait is bounded and traps instead of hanging the GPU.
#pragma once
#include "gnn_tc.cuh"

namespace ldpc {

#ifndef GNN_PIPE_PARTS
#define GNN_PIPE_PARTS 2                // row warps per TMEM lane quarter (each owns 1/PARTS of the columns)
#endif
constexpr int kPipeParts = GNN_PIPE_PARTS;
constexpr int kPipeRowWarps = 4 * kPipeParts, kPipeLoaderWarps = 8;
constexpr int kPipeCw = 64 / kPipeParts, kPipeHw = 128 / kPipeParts;       // comb / output and hidden columns per row thread
constexpr int kPipeMmaWarp = kPipeRowWarps;                                    // warp 8
constexpr int kPipeThreads = (kPipeRowWarps + 1 + kPipeLoaderWarps) * 32;      // 544
constexpr int kPipeRowThreads = kPipeRowWarps * 32, kPipeLoaderThreads = kPipeLoaderWarps * 32;
constexpr size_t kPipeStage = 128 * 64 * sizeof(float);                        // 32 KB
constexpr size_t kPipeSmem = (size_t)(2 * 128 * 64 + 2 * 64 * 128) * sizeof(float) + 3 * kPipeStage;   // 224 KB
constexpr uint32_t kTmAcHi = 0, kTmAcLo = 64, kTmD1 = 128, kTmD2 = 128, kTmHHi = 256, kTmHLo = 384;
// kF16 (fp16 two-way split operands, two k-values per 32-bit column): comb hi [0,32) lo [32,64), h hi [256,320) lo [320,384)
constexpr uint32_t kTmAcLo16 = 32, kTmHLo16 = 320;

enum PipeBar { kBarFullX = 0, kBarFreeX, kBarAReady, kBarD1Full, kBarFullP, kBarFreeP, kBarHReady, kBarD2Full, kBarFreeY,
               kBarD2Drained, kNumPipeBars };

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// split tcgen05.ld: issue now, wait later (tcgen05.wait::ld covers every outstanding load of the thread); the "+r"
// ties keep the compiler from touching the destination registers before the wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&a)[16], uint32_t (&b)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
                   "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]));
    asm volatile("" : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]),
                      "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]));
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// fp16 two-way split (round 2): x = hi + lo, hi = fp16(x), lo = fp16(x - hi) -- 22 mantissa bits; hi.hi + hi.lo + lo.hi on
// kind::f16 measured 4.3e-7 relative on the K = 64 product (3xTF32: 5.1e-7) at TWICE the tensor-pipe rate (773 vs 1 543
// cycles for GEMM1, 1 079 vs 2 176 for GEMM2; tools/probe/umma_probe.cu), and the operands are half as wide.  Values must stay
// below 65 504 (fp16 range); small values lose nothing that matters (lo becomes subnormal below |x| = 0.25: absolute 6e-8).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_gemm3_ts_f16(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K,
                                                   uint32_t sbo_b, uint32_t idesc) {
    uint32_t acc = 0u;
    for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t kb = (uint32_t)ks * 256u;                 // two 16-byte k-chunks (16 halves) of B per MMA; 8 TMEM columns of A
        umma_f16_ts(d_tmem, a_lo + ks * 8, umma_desc(b_hi + kb, 128, sbo_b), idesc, acc);
        umma_f16_ts(d_tmem, a_hi + ks * 8, umma_desc(b_lo + kb, 128, sbo_b), idesc, 1u);
        umma_f16_ts(d_tmem, a_hi + ks * 8, umma_desc(b_hi + kb, 128, sbo_b), idesc, 1u);
        acc = 1u;
    }
}
// 16 consecutive k-values -> 8 packed columns of each image (k even in the low half)
// hi is formed in fp32 with Veltkamp's splitting (t = x * (2^13 + 1); hi = t - (t - x): x rounded to 11 significant bits,
// exactly representable in fp16 for |x| in the normal range), so that only the two packing conversions touch the
// conversion pipe (cvt + unpack + cvt per pair made E1 slower than the TF32 truncation split: 2.5 k vs 1.4 k cycles per tile)
#ifndef GNN_F16_VELTKAMP
#define GNN_F16_VELTKAMP 1
#endif
__device__ __forceinline__ void split16_h(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#if GNN_F16_VELTKAMP
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ta = __fmul_rn(a, 8193.0f), tb = __fmul_rn(b, 8193.0f);
        const float ha = __fsub_rn(ta, __fsub_rn(ta, a)), hb = __fsub_rn(tb, __fsub_rn(tb, b));
        const __half2 h = __floats2half2_rn(ha, hb);
        const __half2 l = __floats2half2_rn(__fsub_rn(a, ha), __fsub_rn(b, hb));
#else
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
#endif
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
// hi = x with the 13 low mantissa bits cleared (a TF32 value), lo = x - hi (exact; the tensor core reads its top 19
// bits).  Truncation instead of cvt.rna costs 2 instructions per element instead of 8 (cvt.rna.tf32.f32 expands to
// FSETP + VIADD + SEL + LOP3 on sm_100) and keeps the 3xTF32 product error at ~2^-20 relative.
#ifndef GNN_LO_ROUND
#define GNN_LO_ROUND 1
#endif
// the tensor core reads only the top 19 bits of lo: scaling lo by (1 + 2^-11) first turns that truncation into
// round-to-nearest (removes the systematic towards-zero bias of the lo term) for one FMUL
constexpr float kLoRound = GNN_LO_ROUND ? 1.00048828125f : 1.0f;
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
        lo[i] = __float_as_uint((v[i] - __uint_as_float(hi[i])) * kLoRound);
    }
}

#ifdef GNN_PIPE_TRACE
#define PIPE_TRACE_DECL long long tr[12][8]; int trn = 0; const bool tracer = blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kPipeMmaWarp || warp == kPipeMmaWarp + 1);
#define PIPE_TRACE(ev) do { if (tracer && trn >= 2 && trn < 14) tr[trn - 2][ev] = clock64(); } while (0)
#define PIPE_TRACE_NEXT() do { ++trn; } while (0)
#define PIPE_TRACE_DUMP(role, nev) do { if (tracer) for (int i = 0; i < 12 && i + 2 < trn; ++i) { printf("%s tile %d:", role, i + 2); \
    for (int e = 0; e < nev; ++e) printf(" %lld", tr[i][e] % 100000000LL); printf("\n"); } } while (0)
#else
#define PIPE_TRACE_DECL
#define PIPE_TRACE(ev)
#define PIPE_TRACE_NEXT()
#define PIPE_TRACE_DUMP(role, nev)
#endif

template <bool kResidual, bool kF16>
__global__ void __launch_bounds__(kPipeThreads, 1) gnn_edge_pipe_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l, const float* __restrict__ tc_l,
    const int* __restrict__ edge_var, const int* __restrict__ edge_chk, const int* __restrict__ edge_type,
    const float* __restrict__ Pv, const float* __restrict__ Pc, long long B, int E, int N, int M, float* __restrict__ y,
    const float* __restrict__ w_out, float* __restrict__ dec_out, int* __restrict__ status) {
    // dec_out != nullptr (last layer, inference): the output projection of the readout (MessageGNNDecoder.decode_messages,
    // message_gnn_decoder.py:46-47) is applied to the finished rows here -- dec_out[row] = <y_row, w_out> -- and y itself is
    // not written: saves the 256-byte row store and its re-read by the readout kernel.
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int kWb = kF16 ? 2 : 4;                       // bytes per weight element (fp16 split images are half as large;
    uint8_t* W1Ahi = tc_smem;                               // [128 x 64]         the staging tiles keep their offsets)
    uint8_t* W1Alo = W1Ahi + 128 * 64 * kWb;
    uint8_t* W2hi = W1Alo + 128 * 64 * kWb;                 // [64 x 128]
    uint8_t* W2lo = W2hi + 64 * 128 * kWb;
    uint8_t* Sx = tc_smem + (size_t)(2 * 128 * 64 + 2 * 64 * 128) * 4;
    // kF16: the half of the weight region the fp16 images leave free holds a staging tile of its OWN for the outputs, so that
    // the loaders may publish x(t+2) while y(t) is still being stored (with 3xTF32 weights y shares Sx and the two serialise:
    // clock64 trace, FullX wait 600-1700 cycles per tile)
    uint8_t* Sy = kF16 ? tc_smem + (size_t)(2 * 128 * 64 + 2 * 64 * 128) * 2 : Sx;
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
        init(kBarFullP, kPipeLoaderThreads); init(kBarFreeP, kPipeRowWarps); init(kBarHReady, kPipeRowWarps); init(kBarD2Full, 1);
        init(kBarFreeY, kPipeRowWarps); init(kBarD2Drained, kPipeRowWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // weights: canonical hi/lo images, contiguous in global memory in the same order as in shared memory
        // (kF16: tc_l points at the layer's fp16 images, kTc16PerLayer halves)
        const float4* src = reinterpret_cast<const float4*>(kF16 ? tc_l : tc_l + kTcW1A);
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < (2 * 128 * 64 + 2 * 64 * 128) * kWb / 16; t += kPipeThreads) dst[t] = src[t];
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
    PIPE_TRACE_DECL
    auto wait = [&](int b, uint32_t parity) { if (ok) ok = mbar_wait(&bars[b], parity & 1u); };
    // one elected lane arrives for the warp once every lane has finished the work the barrier publishes
    auto warp_arrive = [&](int b) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[b]); };

    if (warp < kPipeRowWarps) {
        // ================= row warps: thread <-> message row of the tile =================
        const int rowi = tid & 127, part = tid >> 7;                         // column part of this thread
        const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
        const int cc = tid & 15, cr0 = tid >> 4;                             // cooperative mapping of the y store
        const float4 bias = *reinterpret_cast<const float4*>(b2s + cc * 4);  // b2 of the 4 output columns this thread stores
        // (scalar loads: the flat parameter vector gives w_out no 16-byte alignment)
        const float4 wo = dec_out ? make_float4(__ldg(w_out + cc * 4), __ldg(w_out + cc * 4 + 1), __ldg(w_out + cc * 4 + 2), __ldg(w_out + cc * 4 + 3))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        auto convert = [&]() {                                               // C: Sx -> comb hi/lo in TMEM, my 32 columns
#pragma unroll
            for (int c0 = 0; c0 < kPipeCw; c0 += 16) {
                const int col = part * kPipeCw + c0;
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *stage_ptr(Sx, rowi, (col >> 2) + q);
                    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
                }
                if constexpr (kF16) {
                    uint32_t hi[8], lo[8];
                    split16_h(v, hi, lo);
                    tmem_st8(tmem + lane_base + kTmAcHi + col / 2, hi);
                    tmem_st8(tmem + lane_base + kTmAcLo16 + col / 2, lo);
                } else {
                    uint32_t hi[16], lo[16];
                    split16(v, hi, lo);
                    tmem_st16(tmem + lane_base + kTmAcHi + col, hi);
                    tmem_st16(tmem + lane_base + kTmAcLo + col, lo);
                }
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
            PIPE_TRACE(0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                uint8_t* Sp = part * kPipeHw < 64 ? Sp0 : Sp1;              // hidden columns [0,64): variable-node term, [64,128): check-node term
#pragma unroll
                for (int c0 = 0; c0 < kPipeHw; c0 += 32) {                   // two accumulator loads in flight per wait
                    const int col = part * kPipeHw + c0;
                    const int pch = (col & 63) >> 2;                         // first chunk of these columns in the staged row
                    uint32_t ha[16], hb[16];
                    tmem_ld16_issue(tmem + lane_base + kTmD1 + col, ha);
                    tmem_ld16_issue(tmem + lane_base + kTmD1 + col + 16, hb);
                    float4 pa[4], pb[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { pa[q] = *stage_ptr(Sp, rowi, pch + q); pb[q] = *stage_ptr(Sp, rowi, pch + 4 + q); }
                    tmem_ld_wait(ha, hb);
                    auto finish = [&](uint32_t (&hh)[16], const float4 (&pp)[4], int cofs) {
                        if constexpr (kF16) {
                            float hv[16];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float add[4] = {pp[q].x, pp[q].y, pp[q].z, pp[q].w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) hv[q * 4 + i] = fmaxf(__uint_as_float(hh[q * 4 + i]) + add[i], 0.f);
                            }
                            uint32_t hi[8], lo[8];
                            split16_h(hv, hi, lo);
                            tmem_st8(tmem + lane_base + kTmHHi + (col + cofs) / 2, hi);
                            tmem_st8(tmem + lane_base + kTmHLo16 + (col + cofs) / 2, lo);
                        } else {
                        uint32_t lo[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float add[4] = {pp[q].x, pp[q].y, pp[q].z, pp[q].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float v = fmaxf(__uint_as_float(hh[q * 4 + i]) + add[i], 0.f);
                                hh[q * 4 + i] = __float_as_uint(v) & 0xffffe000u;
                                lo[q * 4 + i] = __float_as_uint((v - __uint_as_float(hh[q * 4 + i])) * kLoRound);
                            }
                        }
                        tmem_st16(tmem + lane_base + kTmHHi + col + cofs, hh);
                        tmem_st16(tmem + lane_base + kTmHLo + col + cofs, lo);
                        }
                    };
                    finish(ha, pa, 0);
                    finish(hb, pb, 16);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(kBarHReady);
            warp_arrive(kBarFreeP);
            PIPE_TRACE(1);
            // C(t+1) while the tensor pipe runs GEMM2(t)
            if (k + 1 < my_tiles) { wait(kBarFullX, (uint32_t)(k + 1)); if (!ok) break; PIPE_TRACE(2); convert(); }
            PIPE_TRACE(3);
            // E2: D2 -> Sx, which C(t+1) has just consumed: each thread overwrites exactly the chunks it read (b2 and the
            // residual are added by the store) (the variable-node rows of this tile are consumed: GEMM2 needed every row warp's h)
            // residual rows requested now (L2 hits: the loaders read them one tile ago), consumed by the store below
            constexpr int kStIters = 128 * 16 / kPipeRowThreads;
            float4 xv[kStIters];
            if constexpr (kResidual) {
#pragma unroll
                for (int it = 0; it < kStIters; ++it) {
                    const long long row = row0 + it * (kPipeRowThreads / 16) + cr0;
                    xv[it] = row < rows ? reinterpret_cast<const float4*>(x + (size_t)row * kH)[cc] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            wait(kBarD2Full, (uint32_t)k);
            if (!ok) break;
            PIPE_TRACE(4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const int col = part * kPipeCw;
                uint32_t oa[16], ob[16];
                tmem_ld16_issue(tmem + lane_base + kTmD2 + col, oa);
                if constexpr (kPipeCw == 32) tmem_ld16_issue(tmem + lane_base + kTmD2 + col + 16, ob);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) ob[i] = 0;
                }
                tmem_ld_wait(oa, ob);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    *stage_ptr(Sy, rowi, (col >> 2) + q) = make_float4(__uint_as_float(oa[q * 4]), __uint_as_float(oa[q * 4 + 1]),
                                                                        __uint_as_float(oa[q * 4 + 2]), __uint_as_float(oa[q * 4 + 3]));
                    if constexpr (kPipeCw == 32)
                        *stage_ptr(Sy, rowi, (col >> 2) + 4 + q) = make_float4(__uint_as_float(ob[q * 4]), __uint_as_float(ob[q * 4 + 1]),
                                                                                __uint_as_float(ob[q * 4 + 2]), __uint_as_float(ob[q * 4 + 3]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(kBarD2Drained);
            PIPE_TRACE(5);
            asm volatile("bar.sync 1, %0;" :: "n"(kPipeRowThreads) : "memory");
            // ST: coalesced rows (+ residual x)
#pragma unroll
            for (int it = 0; it < kStIters; ++it) {
                const int rr = it * (kPipeRowThreads / 16) + cr0;
                float4 r = *stage_ptr(Sy, rr, cc);
                r.x += bias.x; r.y += bias.y; r.z += bias.z; r.w += bias.w;
                if constexpr (kResidual) { r.x += xv[it].x; r.y += xv[it].y; r.z += xv[it].z; r.w += xv[it].w; }
                if (dec_out) {                                           // warp-uniform
                    float d = __fmaf_rn(r.w, wo.w, __fmaf_rn(r.z, wo.z, __fmaf_rn(r.y, wo.y, r.x * wo.x)));
#pragma unroll
                    for (int sft = 8; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);   // the 16 lanes of this row
                    if (cc == 0 && row0 + rr < rows) dec_out[row0 + rr] = d;
                } else if (row0 + rr < rows) {
                    reinterpret_cast<float4*>(y + (size_t)(row0 + rr) * kH)[cc] = r;
                }
            }
            warp_arrive(kBarFreeY);
            PIPE_TRACE(6);
            PIPE_TRACE_NEXT();
        }
        PIPE_TRACE_DUMP("row", 7);
    } else if (warp == kPipeMmaWarp) {
        // ================= MMA issue (one thread) =================
        if (lane == 0) {
            constexpr uint32_t kIdesc128 = kF16 ? umma_idesc_f16(128) : umma_idesc_tf32(128), kIdesc64 = kF16 ? umma_idesc_f16(64) : umma_idesc_tf32(64);
            auto gemm1 = [&]() {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (kF16) umma_gemm3_ts_f16(tmem + kTmD1, tmem + kTmAcHi, tmem + kTmAcLo16, smem_u32(W1Ahi), smem_u32(W1Alo), 64, 1024, kIdesc128);
                else umma_gemm3_ts(tmem + kTmD1, tmem + kTmAcHi, tmem + kTmAcLo, smem_u32(W1Ahi), smem_u32(W1Alo), 64, 2048, kIdesc128);
                umma_commit(&bars[kBarD1Full]);
            };
            if (my_tiles > 0) { wait(kBarAReady, 0); if (ok) gemm1(); }
            for (long long k = 0; k < my_tiles && ok; ++k) {
                wait(kBarHReady, (uint32_t)k);
                if (!ok) break;
                PIPE_TRACE(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (kF16) umma_gemm3_ts_f16(tmem + kTmD2, tmem + kTmHHi, tmem + kTmHLo16, smem_u32(W2hi), smem_u32(W2lo), 128, 2048, kIdesc64);
                else umma_gemm3_ts(tmem + kTmD2, tmem + kTmHHi, tmem + kTmHLo, smem_u32(W2hi), smem_u32(W2lo), 128, 4096, kIdesc64);
                umma_commit(&bars[kBarD2Full]);
                PIPE_TRACE(1);
                if (k + 1 < my_tiles) {
                    wait(kBarAReady, (uint32_t)(k + 1));
                    wait(kBarD2Drained, (uint32_t)k);                     // D2(t) shares columns with D1
                    if (!ok) break;
                    PIPE_TRACE(2);
                    gemm1();
                    PIPE_TRACE(3);
                }
                PIPE_TRACE_NEXT();
            }
            PIPE_TRACE_DUMP("mma", 4);
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
        float4 X[2][4];
        int ty[2], vrow[2], crow[2];                                          // message type; Pv / Pc row (-1: past the end)
        int ev[2], ec[2], vb[2], cb[2];                                       // raw indices of the tile in flight
        auto load_x = [&](long long k, const Pos& p) {                        // x rows + all indices of tile k (requests only)
            const long long row0 = (blockIdx.x + k * gridDim.x) * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = lrow + 64 * h;
                int e = p.e0 + rr;
                long long b = p.b0;
                while (e >= E) { e -= E; ++b; }
                const bool valid = row0 + rr < rows;
                const float4* src = reinterpret_cast<const float4*>(x + (size_t)(row0 + rr) * kH);
#pragma unroll
                for (int j = 0; j < 4; ++j) X[h][j] = valid ? src[4 * j + q4] : make_float4(0.f, 0.f, 0.f, 0.f);
                ty[h] = valid ? __ldg(edge_type + e) : 0;
                ev[h] = valid ? __ldg(edge_var + e) : 0;
                ec[h] = valid ? __ldg(edge_chk + e) : 0;
                vb[h] = valid ? (int)(b * N) : -1;
                cb[h] = valid ? (int)(b * M) : -1;
            }
        };
        auto take_indices = [&]() {                                           // first use of the index requests
#pragma unroll
            for (int h = 0; h < 2; ++h) { vrow[h] = vb[h] < 0 ? -1 : vb[h] + ev[h]; crow[h] = cb[h] < 0 ? -1 : cb[h] + ec[h]; }
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
        if (my_tiles > 0) { load_x(0, pos); take_indices(); publish_x(); }
        for (long long k = 0; k < my_tiles && ok; ++k) {
            const bool more = k + 1 < my_tiles;
            advance(pos);
            if (more) load_x(k + 1, pos);                                     // in flight during everything below
            // node rows of tile k: asynchronous copies straight into Sp0 / Sp1 once E1(k-1) has released them; the
            // copies signal the row warps themselves (cp.async.mbarrier.arrive), the loader does not wait for them
            wait(kBarFreeP, (uint32_t)k + 1u);
            if (!ok) break;
            PIPE_TRACE(0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* sv = Pv + (size_t)(vrow[h] < 0 ? 0 : vrow[h]) * kH;
                const float* sc = Pc + (size_t)(crow[h] < 0 ? 0 : crow[h]) * kH;
                const int nbytes = vrow[h] >= 0 ? 16 : 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(stage_ptr(Sp0, lrow + 64 * h, 4 * j + q4))),
                                 "l"(sv + (4 * j + q4) * 4), "r"(nbytes) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(stage_ptr(Sp1, lrow + 64 * h, 4 * j + q4))),
                                 "l"(sc + (4 * j + q4) * 4), "r"(nbytes) : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&bars[kBarFullP])) : "memory");
            if (more) {
                take_indices();                                               // of tile k+1, requested at the top
                wait(kBarFreeX, (uint32_t)k);                                      // C(k) has read Sx ...
                if constexpr (!kF16) wait(kBarFreeY, (uint32_t)k + 1u);             // ... and y(k-1) has left it (own tile with kF16)
                if (!ok) break;
                PIPE_TRACE(1);
                publish_x();
                PIPE_TRACE(2);
            }
            PIPE_TRACE_NEXT();
        }
        PIPE_TRACE_DUMP("load", 3);
    }
    if (!ok) { atomicExch(status, 1); asm volatile("trap;"); }               // a hand-off never arrived: fail loudly
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();
    if (warp == kPipeMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ---- node kernel on tensor cores: P[b][node][:] = W1B . mean_{e in node}(x + emb) + b1 ----------------
// 512 threads per 128-node tile.  Cooperative phases (16 threads per node, one 16-byte chunk each) do all the
// global traffic as coalesced 256-byte rows: gather-and-average the node's message rows, and store the P rows;
// per-row phases (4 threads per node) split the means hi/lo into TENSOR MEMORY (the A operand, as in the edge
// kernel) and read the accumulator.  Without an activation tile in shared memory the CTA needs 64 KB (W1B hi/lo +
// staging) and 256 TMEM columns, so TWO CTAs are resident per SM and one CTA's gather (latency-bound: a dependent
// walk over the node's edge list) overlaps the other's MMA, epilogue and store.
constexpr int kNodeParts = 4;
constexpr int kNodeThreads = 128 * kNodeParts;
constexpr size_t kNodeTcSmem = (size_t)(2 * 64 * 64) * sizeof(float) + 128 * 64 * sizeof(float);   // 32 + 32 KB
constexpr uint32_t kTmNodeAHi = 0, kTmNodeALo = 64, kTmNodeD = 128;

__global__ void __launch_bounds__(kNodeThreads, 2) gnn_node_tc_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l, const float* __restrict__ tc_l,
    int kind, const int* __restrict__ ptr, const int* __restrict__ list, const int* __restrict__ edge_type, long long B, int E,
    int nodes, float* __restrict__ P, float* __restrict__ Msave, int* __restrict__ status) {
    // Msave != nullptr (training forward): the node means are kept for the backward pass (weight gradient of W1B)
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* Whi = tc_smem;                                 // W1B [64 x 64]
    uint8_t* Wlo = Whi + 64 * 64 * 4;
    uint8_t* S = Wlo + 64 * 64 * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float b1s[kH];
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
        const float4* src = reinterpret_cast<const float4*>(tc_l + (kind == 0 ? kTcW1BV : kTcW1BC));
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < 2 * 64 * 64 / 4; t += kNodeThreads) dst[t] = src[t];
        if (tid < kH) b1s[tid] = packed_l[(kind == 0 ? kPkB1V : kPkB1C) + tid];
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
    const long long rows = B * nodes, tiles = (rows + 127) / 128;
    const int cc = tid & 15, cr0 = tid >> 4;                 // cooperative mapping: chunk, first node of the pass
    for (long long tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const long long row0 = tile * 128;
        // 0. cooperative: mean over the node's messages of (x + emb), chunk cc -> S
#pragma unroll 1
        for (int it = 0; it < 128 * 16 / kNodeThreads; ++it) {
            const int rr = it * (kNodeThreads / 16) + cr0;
            const long long row = row0 + rr;
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows) {
                const int node = (int)(row % nodes);
                const long long b = row / nodes;
                const int k0 = ptr[node], k1 = ptr[node + 1];
                const float* xb = x + (size_t)b * E * kH;
                int q = k0;
                for (; q + 4 <= k1; q += 4) {                 // four message rows in flight per thread
                    int e[4];
                    float4 a[4], em[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = list ? list[q + u] : q + u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        a[u] = reinterpret_cast<const float4*>(xb + (size_t)e[u] * kH)[cc];
                        em[u] = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[e[u]] * kH) + cc);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { m.x += a[u].x + em[u].x; m.y += a[u].y + em[u].y; m.z += a[u].z + em[u].z; m.w += a[u].w + em[u].w; }
                }
                for (; q < k1; ++q) {
                    const int e = list ? list[q] : q;
                    const float4 a = reinterpret_cast<const float4*>(xb + (size_t)e * kH)[cc];
                    const float4 em = __ldg(reinterpret_cast<const float4*>(emb_l + (size_t)edge_type[e] * kH) + cc);
                    m.x += a.x + em.x; m.y += a.y + em.y; m.z += a.z + em.z; m.w += a.w + em.w;
                }
                const float inv = 1.0f / (float)(k1 - k0);
                m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
                if (Msave) reinterpret_cast<float4*>(Msave + (size_t)row * kH)[cc] = m;
            }
            *stage_ptr(S, rr, cc) = m;
        }
        __syncthreads();
        // 1. per row: S -> split hi/lo -> Tensor Memory (A operand)
        {
            constexpr int kCols = 64 / kNodeParts;               // 16
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = *stage_ptr(S, rowi, part * (kCols / 4) + q);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
            uint32_t hi[16], lo[16];
            split16(v, hi, lo);
            tmem_st16(tmem + my_lane + kTmNodeAHi + part * kCols, hi);
            tmem_st16(tmem + my_lane + kTmNodeALo + part * kCols, lo);
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
        // 2. per row: D + b1 -> S;  cooperative: S -> P (coalesced rows)
        if (ok) {
#pragma unroll
            for (int c0 = 0; c0 < 64 / kNodeParts; c0 += 16) {
                const int col = part * (64 / kNodeParts) + c0;
                float o[16];
                tmem_ld16(tmem + my_lane + kTmNodeD + col, o);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *stage_ptr(S, rowi, (col >> 2) + q) = make_float4(o[q * 4] + b1s[col + q * 4], o[q * 4 + 1] + b1s[col + q * 4 + 1],
                                                                      o[q * 4 + 2] + b1s[col + q * 4 + 2], o[q * 4 + 3] + b1s[col + q * 4 + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
#pragma unroll
            for (int it = 0; it < 128 * 16 / kNodeThreads; ++it) {
                const int rr = it * (kNodeThreads / 16) + cr0;
                if (row0 + rr < rows) reinterpret_cast<float4*>(P + (size_t)(row0 + rr) * kH)[cc] = *stage_ptr(S, rr, cc);
            }
        }
        __syncthreads();
    }
    if (!ok) { if (tid == 0) atomicExch(status, 1); asm volatile("trap;"); }     // MMA completion never arrived: fail loudly
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}


}  // namespace ldpc
