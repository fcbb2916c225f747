// decode_fast.cuh -- scaled min-sum, specialised at compile time for the shipped 5G tables
// (NR_2_0_32 and NR_2_0_4: BG2, set index 0, Z = 32 / 4).  LDPC_PATH_FAST.
//
// Replaces the hot loop of MinSumScaledDecoder.decode (models/traditional_decoders.py:205-252)
// for the headline workload (10 iterations, BG2 Z=32).  Design, B200-first:
//   * one warp <-> 32/Z codewords, lane <-> row r of every Z x Z circulant; the whole base
//     graph (42 rows, 197 cells) is unrolled from constexpr tables, so every register index,
//     shared-memory offset and shuffle distance is an immediate;
//   * posteriors T[j] of the 14 "core" columns (degree > 1) live in registers,
//     VARIABLE-aligned; a circulant shift s is one __shfl_sync rotation (lane+s) into check
//     alignment, and (lane-s) back for the accumulation (sub-warp width=4 groups at Z=4);
//   * the 38 degree-1 columns never change their variable-to-check message (it is always the
//     channel LLR, reference :235-244 with an empty sum), so they stay in registers, need no
//     rotation (their shift is 0), no message storage, and their posterior is only formed in
//     the last iteration;
//   * check-to-variable messages of the 159 core cells stay resident in shared memory for all
//     iterations as per-lane private quads (float4 mq[q][lane]: conflict-free 128-bit LDS/STS,
//     no barriers, a quarter of the load/store instructions);
//   * the iteration body is one branch-free basic block (the last iteration is a second
//     instantiation), so ptxas can interleave independent check rows to cover the ~30-cycle
//     shared-memory/shuffle latency with only 10 resident warps per SM;
//   * check node: running min1/min2 + XOR of sign bits; the message to edge k is
//     alpha*(|v_k|==min1 ? min2 : min1) with sign (total ^ sign_k) -- identical to the
//     reference's product-of-signs / min-over-others including its sign(0)=0 rule, because a
//     zero input is the minimum and zeroes every other output by magnitude;
//   * variable node: v2c = T - c2v (total minus self) instead of the reference's sum over the
//     other checks: this is the one place the operation order differs (<= a few ulp of T per
//     iteration; the posterior itself is accumulated in the reference's ascending-check
//     order).  LDPC_PATH_EXACT keeps the reference order bit for bit.
#pragma once
#include "math_ref.cuh"
#include <math_constants.h>

#include "bg2_tables.h"
#include "params.cuh"

namespace ldpc {

// Register budget: warps are spread over the 4 SM sub-partitions (16K registers each), so a
// 10-warp CTA has sub-partitions with 3 warps -> at most 16384/96 = 170 registers per thread;
// __launch_bounds__ lets ptxas derive that cap (168) itself.
// ---- sum-product helpers (kAlgo == LDPC_ALGO_BP) ---------------------------------------------------
// The reference's variable update adds the OTHER checks' messages in order, unclipped, so +-inf and
// NaN messages are routine and "posterior minus own message" would give inf-inf.
// Per column the kernel therefore keeps F = llr + sum of the FINITE messages (ascending check
// order) and K = count of +inf messages (bits 0-7) | count of -inf messages (bits 8-15); a NaN message
// counts as both (either way every OTHER edge of the column sees NaN).  The sum over the others is
// rebuilt from (F - own finite part, K - own code) with IEEE semantics.
// BP messages are stored VARIABLE-aligned (lane <-> variable row of the circulant), so F, K and the own
// message meet in the same lane: only the resolved v2c and the new c2v cross lanes (2 rotations per edge).
struct BpCode {
    int code;       // contribution to K, 0 for a finite message
    bool nonfin;
};
__device__ __forceinline__ BpCode bp_code(float x) {
    BpCode c;
    c.nonfin = !(fabsf(x) < CUDART_INF_F);
    int k = (int)(f2u(x) >> 31) * 0xff + 1;          // +inf -> 0x001, -inf -> 0x100
    k = (x != x) ? 0x101 : k;
    c.code = c.nonfin ? k : 0;
    return c;
}
__device__ __forceinline__ float bp_resolve(float finite_sum, int k) {
    float r = (k & 0xff) ? CUDART_INF_F : finite_sum;
    r = (k & 0xff00) ? __fadd_rn(r, -CUDART_INF_F) : r;          // +inf and -inf present: NaN
    return r;
}

// tanh(v/2) and 2*atanh(p) on the special-function unit, inline (7 instructions each, 2 MUFU):
//   tanh(v/2)  = 1 - 2e/(1+e),  e = 2^(-|v| log2 e)      -- one rounding in the final FMA, so the result saturates to
//                exactly 1.0f where the correctly rounded tanh does (e < 2^-26, |v| > 18.02) up to the 2-ulp error of
//                ex2.approx; v = +-inf gives e = 0 -> 1, NaN propagates;
//   2 atanh(p) = ln((1+|p|)/(1-|p|)) = lg2(q) ln 2        -- |p| = 1 gives rcp(0) = inf -> inf, NaN propagates.
// Absolute error of both is ~1e-7 (relative error grows for |v|, |p| << 1, where the message is added to O(1) terms);
// the previous out-of-line tanhf/atanhf cost ~55 instructions + two calls per edge.
// LDPC_BP_REF_FN=1: the exact kernel's once-rounded double evaluation (math_ref.cuh), out of line (diagnostics).
#ifndef LDPC_BP_REF_FN
#define LDPC_BP_REF_FN 0
#endif
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#if LDPC_BP_REF_FN
__device__ __noinline__ float bp_tanh_half(float v) { return tanh_ref(0.5f * v); }
__device__ __noinline__ float bp_two_atanh(float p) { return 2.0f * atanh_ref(p); }
#else
__device__ __forceinline__ float bp_tanh_half(float v) {
    const float e = mufu_ex2(__fmul_rn(fabsf(v), -1.4426950408889634f));
    const float r = mufu_rcp(__fadd_rn(1.0f, e));
    const float t = __fmaf_rn(__fmul_rn(e, r), -2.0f, 1.0f);
    return u2f(f2u(t) | (f2u(v) & 0x80000000u));
}
__device__ __forceinline__ float bp_two_atanh(float p) {
    const float a = fabsf(p);
    const float q = __fmul_rn(__fadd_rn(1.0f, a), mufu_rcp(__fsub_rn(1.0f, a)));
    const float r = __fmul_rn(mufu_lg2(q), 0.6931471805599453f);
    return u2f(f2u(r) | (f2u(p) & 0x80000000u));
}
#endif

// ---- optional: resident messages in Tensor Memory instead of shared memory (LDPC_FAST_TMEM=1) -------
// TMEM is 128 lanes x 512 columns x 32 bit per SM; a warp reaches the 32 lanes of its quarter
// (warp_id % 4) with tcgen05.ld/st.32x32b, i.e. exactly a per-lane private array: lane i <-> TMEM
// lane, message quad q <-> columns 4q..4q+3.  Moving the 160 message words there takes 320 of the
// 622 wavefronts per codeword-iteration off the shared-memory/shuffle pipe (the binding unit).
#ifndef LDPC_FAST_TMEM
#define LDPC_FAST_TMEM 2
#endif
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
    // the wait carries the loaded registers so that no consumer is scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d));
    return make_float4(u2f(a), u2f(b), u2f(c), u2f(d));
}
// split form for software pipelining: issue now, settle (wait + re-define the registers) later
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, float& a, float& b, float& c, float& d) {
    uint32_t x, y, z, w;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr));
    a = u2f(x); b = u2f(y); c = u2f(z); d = u2f(w);
}
__device__ __forceinline__ void tmem_tie4(float& a, float& b, float& c, float& d) {
    // empty volatile asm after the wait: consumers depend on these re-definitions, so none of them
    // can be scheduled above the tcgen05.wait::ld that precedes it
    asm volatile("" : "+f"(a), "+f"(b), "+f"(c), "+f"(d));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(f2u(v.x)), "r"(f2u(v.y)), "r"(f2u(v.z)), "r"(f2u(v.w)) : "memory");
}

// Four channel LLRs (all-zero codeword) of one Philox block.  Out of line on purpose: inlined 13 times the
// generator is ~80 KB of SASS in front of the iteration body and evicts it from the instruction cache once per
// codeword (the fused simulation ran at half the decode-only rate).
__device__ __noinline__ float4 gen_llr4(const GenParams g, unsigned long long frame, uint32_t blk) {
    float z[4];
    normal4(g.seed, frame, blk, z);
    return make_float4(llr_from_noise(z[0], 1.0f, g), llr_from_noise(z[1], 1.0f, g), llr_from_noise(z[2], 1.0f, g),
                       llr_from_noise(z[3], 1.0f, g));
}

template <class BG>
constexpr size_t fast_smem_bytes(int warps) {
    return (size_t)warps * ((LDPC_FAST_TMEM ? 0 : (BG::kCoreEdges + 3) / 4) + (BG::kCoreCols + 3) / 4) * 32 * sizeof(float4);
}

// kEarly: per-codeword early exit (LDPC_STOP_PER_CODEWORD).  With Z < 32 a warp carries 32/Z codewords: each FREEZES its
// decisions and iteration count at its own first valid iteration, the warp runs until all of them have.  Every iteration then also
// forms the hard decisions of the degree-1 columns and the syndrome of the new posteriors; a codeword stops after its
// first valid iteration (traditional_decoders.py:102-106 / 255-258 applied per codeword, as the exact kernel does).
// Outputs in this mode: hard decisions, syndrome_ok, iters_out (soft_out is served by the exact kernel).
template <class BG, int kWarps, int kAlgo, bool kEarly = false>
__global__ void __launch_bounds__(kWarps * 32, 1) decode_fast_kernel(const DecodeParams p) {
    constexpr int Z = BG::kZ, G = 32 / Z, NC = BG::kCoreCols, NX = BG::kExtCols, EC = BG::kCoreEdges;
    constexpr int EQ = (EC + 3) / 4;                      // message quads per lane
    constexpr int N = BG::kCols * Z, NW = (N + 31) / 32, NWR = (NW + Z - 1) / Z;
    static_assert(NX <= 64, "degree-1 column bitmap is 64 bits");
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-lane private storage: quad q of this lane at mq[q*32]; conflict-free 128-bit accesses
    constexpr int kMsgQuads = LDPC_FAST_TMEM ? 0 : EQ;    // messages live in TMEM when LDPC_FAST_TMEM != 0
    float4* mq = smem4 + warp * ((kMsgQuads + (NC + 3) / 4) * 32) + lane;
    float* Ls = reinterpret_cast<float*>(smem4 + warp * ((kMsgQuads + (NC + 3) / 4) * 32) + kMsgQuads * 32) + lane;   // Ls[k*32]
    (void)mq;
    const int cwi = lane / Z, r = lane % Z;
    const float alpha = p.alpha;
#if LDPC_FAST_TMEM
    static_assert(((kWarps + 3) / 4) * EQ * 4 <= 512, "warps sharing a TMEM lane quarter each need EQ*4 columns");
    __shared__ uint32_t tmem_base_s;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tq = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)(EQ * 4);   // this warp's window
#endif
    unsigned long long acc_bits = 0, acc_fe = 0, acc_frames = 0, acc_und = 0;
    // source-lane operands of the rotations, lane+t for t = 1..Z-1, kept in registers (opaque to
    // the compiler so they are not recomputed with an integer add in front of every shuffle)
#ifndef LDPC_FAST_LP
#define LDPC_FAST_LP 1
#endif
#ifndef LDPC_FAST_FENCE
#define LDPC_FAST_FENCE 2
#endif
    int lp[Z];
    static_for<0, Z>([&](auto tc) {
        constexpr int t = decltype(tc)::value;
#if LDPC_FAST_LP
        asm("add.s32 %0, %1, %2;" : "=r"(lp[t]) : "r"(lane), "n"(t));
#else
        lp[t] = lane + t;
#endif
    });

    for (long long grp = (long long)blockIdx.x * kWarps + warp; grp < p.ngroups; grp += (long long)gridDim.x * kWarps) {
        const long long cw = grp * G + cwi;
        const bool live = cw < p.B;
        float Tc[NC], Lx[NX];
        int Kc[NC];                    // BP only: non-finite message counts per core column
        static_for<0, NC>([&](auto kc) { Kc[decltype(kc)::value] = 0; });
        auto put_llr = [&](auto jc, float x) {
            constexpr int j = decltype(jc)::value;
            constexpr int sl = BG::col_slot[j];
            if constexpr (BG::col_kind[j] == 0) {
                Tc[sl] = x;
                Ls[sl * 32] = x;
            } else {
                Lx[sl] = x;
            }
        };
        if (p.gen.enabled) {
            // on-chip channel: one Philox block per lane yields four 32-wide columns (channel.cuh)
            constexpr int P = 32 / Z, HI = (BG::kCols + 4 * P - 1) / (4 * P);
            const unsigned long long frame = p.gen.first_frame + (unsigned long long)cw;
            static_for<0, HI>([&](auto hc) {
                static_for<0, P>([&](auto mc) {
                    constexpr int hi = decltype(hc)::value, jm = decltype(mc)::value;
                    if constexpr (hi * 4 * P + jm < BG::kCols) {
                        const float4 q = gen_llr4(p.gen, frame, (uint32_t)(hi * 32 + jm * Z + r));
                        const float z[4] = {q.x, q.y, q.z, q.w};
                        static_for<0, 4>([&](auto cc) {
                            constexpr int j = (hi * 4 + decltype(cc)::value) * P + jm;
                            if constexpr (j < BG::kCols) put_llr(IC<j>{}, live ? z[decltype(cc)::value] : 0.0f);
                        });
                    }
                });
            });
        } else {
            const float* llr = p.llr + cw * N + r;
            static_for<0, BG::kCols>([&](auto jc) {
                put_llr(jc, live ? __ldg(llr + decltype(jc)::value * Z) : 0.0f);
            });
            // pull the LLR rows of this warp's NEXT group into L2 while this one is decoded
            const long long ngrp = grp + (long long)gridDim.x * kWarps;
            if (ngrp < p.ngroups) {
                const char* nxt = reinterpret_cast<const char*>(p.llr + ngrp * G * N);
                const long long bytes = (((ngrp + 1) * G <= p.B) ? (long long)G : (p.B - ngrp * G)) * N * (long long)sizeof(float);
#pragma unroll
                for (int q = 0; q < (G * N * 4 + 32 * 128 - 1) / (32 * 128); ++q) {
                    const long long off = ((long long)q * 32 + lane) * 128;
                    if (off < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
                }
            }
        }
#if LDPC_FAST_TMEM
        static_for<0, EQ>([&](auto qc) { tmem_st4(tq + decltype(qc)::value * 4, make_float4(0.f, 0.f, 0.f, 0.f)); });
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#else
        static_for<0, EQ>([&](auto qc) { mq[decltype(qc)::value * 32] = make_float4(0.f, 0.f, 0.f, 0.f); });
#endif

        // One flooding iteration, fully unrolled and branch-free.  kLast additionally forms the
        // posteriors of the degree-1 columns (in place of their channel LLR registers).
        float ov[EQ * 4];                  // SSA view of the old message quads (registers, short-lived)
#if LDPC_FAST_TMEM == 2
        // software-pipelined TMEM loads: the quads first touched by row group g+1 are issued when
        // group g starts, and settled (one tcgen05.wait::ld) when group g+1 starts
        constexpr int kGrp = LDPC_FAST_FENCE, kNumGrp = (BG::kRows + kGrp - 1) / kGrp;
        auto for_group_quads = [&](auto gc, auto&& fn) {
            constexpr int g = decltype(gc)::value;
            constexpr int ea = BG::row_ptr[g * kGrp], eb = BG::row_ptr[(g + 1) * kGrp < BG::kRows ? (g + 1) * kGrp : BG::kRows];
            static_for<ea, eb>([&](auto ec) {
                constexpr int e = decltype(ec)::value;
                if constexpr (BG::kind[e] == 0) {
                    constexpr int mi = BG::msg[e];
                    if constexpr (mi % 4 == 0) fn(IC<mi>{});
                }
            });
        };
        auto issue_group = [&](auto gc) {
            for_group_quads(gc, [&](auto mc) {
                constexpr int mi = decltype(mc)::value;
                tmem_ld4_issue(tq + mi, ov[mi], ov[mi + 1], ov[mi + 2], ov[mi + 3]);
            });
        };
        auto settle_group = [&](auto gc) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for_group_quads(gc, [&](auto mc) {
                constexpr int mi = decltype(mc)::value;
                tmem_tie4(ov[mi], ov[mi + 1], ov[mi + 2], ov[mi + 3]);
            });
        };
        issue_group(IC<0>{});
#endif
        unsigned long long xneg_it = 0;        // kEarly: hard decisions of the degree-1 columns after this iteration
        auto iteration = [&](auto lastc) {
            // mode 0: plain; 1: last (posteriors of the degree-1 columns replace their LLR registers);
            // 2: checked (kEarly): degree-1 hard decisions into xneg_it, LLR registers untouched
            constexpr int kMode = decltype(lastc)::value;
            constexpr bool kLast = kMode != 0;          // "compute the messages towards the degree-1 columns as well"
            float Tn[NC];
            int Kn[NC];
            float nv[EQ * 4];
            static_for<0, NC>([&](auto kc) {
                Tn[decltype(kc)::value] = Ls[decltype(kc)::value * 32];
                Kn[decltype(kc)::value] = 0;
            });
            static_for<0, BG::kRows>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int e0 = BG::row_ptr[i], d = BG::row_ptr[i + 1] - e0;
#if LDPC_FAST_FENCE
                if constexpr (i % LDPC_FAST_FENCE == 0 && i > 0) asm volatile("" ::: "memory");
#endif
#if LDPC_FAST_TMEM == 2
                if constexpr (i % kGrp == 0) {
                    settle_group(IC<i / kGrp>{});
                    if constexpr (i / kGrp + 1 < kNumGrp) issue_group(IC<i / kGrp + 1>{});
                }
#endif
                float v[d];
                // gather variable-to-check messages, check-aligned: v = T - c2v (core), llr (degree-1)
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int c = BG::slot[e], s = BG::shift[e], mi = BG::msg[e];
                        if constexpr (mi % 4 == 0) {
#if LDPC_FAST_TMEM == 1
                            const float4 q4 = tmem_ld4(tq + (mi / 4) * 4);
                            ov[mi] = q4.x; ov[mi + 1] = q4.y; ov[mi + 2] = q4.z; ov[mi + 3] = q4.w;
#elif LDPC_FAST_TMEM == 0
                            const float4 q4 = mq[(mi / 4) * 32];
                            ov[mi] = q4.x; ov[mi + 1] = q4.y; ov[mi + 2] = q4.z; ov[mi + 3] = q4.w;
#endif
                        }
                        if constexpr (kAlgo == LDPC_ALGO_MINSUM) {
                            const float t = (s == 0) ? Tc[c] : __shfl_sync(kFull, Tc[c], lp[s], Z);
                            v[k] = t - ov[mi];
                        } else {
                            // variable-aligned: F, K and the own message are in this lane; rotate the resolved v2c
                            const BpCode own = bp_code(ov[mi]);
                            const float f = own.nonfin ? Tc[c] : __fsub_rn(Tc[c], ov[mi]);
                            const float vv = bp_resolve(f, Kc[c] - own.code);
                            v[k] = (s == 0) ? vv : __shfl_sync(kFull, vv, lp[s], Z);
                        }
                    } else {
                        static_assert(BG::kind[e] == 0 || BG::shift[e] == 0, "degree-1 columns are expected unshifted");
                        constexpr int x = BG::slot[e];
                        v[k] = Lx[x];
                    }
                });
                float cn[d];                       // new check-to-variable messages of this row
                if constexpr (kAlgo == LDPC_ALGO_MINSUM) {
                    // two smallest magnitudes and the sign parity
                    float m1 = fabsf(v[0]), m2 = CUDART_INF_F;
                    unsigned sg = f2u(v[0]);
                    static_for<1, d>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        const float a = fabsf(v[k]);
                        m2 = (k == 1) ? fmaxf(m1, a) : fminf(m2, fmaxf(m1, a));
                        m1 = fminf(m1, a);
                        sg ^= f2u(v[k]);
                    });
#ifndef LDPC_FAST_IMAD_SELECT
#define LDPC_FAST_IMAD_SELECT 1
#endif
#if LDPC_FAST_IMAD_SELECT
                    // ALU-pipe diet (the binding unit): the two candidate messages differ by an integer
                    // constant per row, so the select is an integer multiply-add on the FMA pipe:
                    //   c = p1 ^ (v & signbit)                 LOP3        (ALU)   [xor of the top bit == add mod 2^32]
                    //   P = (|v| == m1)                        FSETP       (ALU)
                    //   @P c += p2 - p1                        IMAD.IADD   (FMA pipe, predicated)
                    const unsigned q1 = f2u(__fmul_rn(alpha, m1)), q2 = f2u(__fmul_rn(alpha, m2));
                    const unsigned p1 = q1 ^ (sg & 0x80000000u);
                    const int ndp = (int)q1 - (int)q2;
                    static_for<0, d>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        if constexpr (BG::kind[e0 + k] == 0 || kLast) {
                            unsigned c = p1 ^ (f2u(v[k]) & 0x80000000u);
                            asm("{ .reg .pred p; setp.eq.f32 p, %1, %2; @p mad.lo.s32 %0, %3, 1, %0; }"
                                : "+r"(c) : "f"(fabsf(v[k])), "f"(m1), "r"(-ndp));
                            cn[k] = u2f(c);
                        }
                    });
#else
                    sg &= 0x80000000u;
                    const unsigned p1 = f2u(__fmul_rn(alpha, m1)) ^ sg, p2 = f2u(__fmul_rn(alpha, m2)) ^ sg;
                    static_for<0, d>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        if constexpr (BG::kind[e0 + k] == 0 || kLast) {
                            const unsigned sel = (fabsf(v[k]) == m1) ? p2 : p1;
                            cn[k] = u2f(sel ^ (f2u(v[k]) & 0x80000000u));
                        }
                    });
#endif
                } else {
                    // sum-product: product of tanh(v/2) over the OTHER edges in ascending order (running
                    // prefix x suffix chain = the reference's multiplication order), 2*atanh, unclipped
                    float t[d];
                    static_for<0, d>([&](auto kc) { t[decltype(kc)::value] = bp_tanh_half(v[decltype(kc)::value]); });
                    float pre = 1.0f;
                    static_for<0, d>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        if constexpr (BG::kind[e0 + k] == 0 || kLast) {
                            float pr = pre;
                            static_for<k + 1, d>([&](auto k2) { pr = __fmul_rn(pr, t[decltype(k2)::value]); });
                            cn[k] = bp_two_atanh(pr);
                        }
                        pre = __fmul_rn(pre, t[k]);
                    });
                }
                // store the messages, accumulate the new posteriors in ascending row order
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int c = BG::slot[e], s = BG::shift[e], mi = BG::msg[e];
                        const float cnew = cn[k];
                        // min-sum stores the message check-aligned, BP variable-aligned (rotated back first)
                        const float tback = (s == 0) ? cnew : __shfl_sync(kFull, cnew, lp[Z - s], Z);
                        nv[mi] = (kAlgo == LDPC_ALGO_MINSUM) ? cnew : tback;
                        if constexpr (mi % 4 == 3 || mi == EC - 1) {
                            constexpr int b = (mi / 4) * 4;
                            const float4 q4 = make_float4(nv[b], b + 1 < EC ? nv[b + 1] : 0.f, b + 2 < EC ? nv[b + 2] : 0.f,
                                                          b + 3 < EC ? nv[b + 3] : 0.f);
#if LDPC_FAST_TMEM
                            tmem_st4(tq + (mi / 4) * 4, q4);
#else
                            mq[(mi / 4) * 32] = q4;
#endif
                        }
                        const float t = tback;
                        if constexpr (kAlgo == LDPC_ALGO_MINSUM) {
                            Tn[c] = __fadd_rn(Tn[c], t);
                        } else {
                            const BpCode cd = bp_code(t);
                            Tn[c] = cd.nonfin ? Tn[c] : __fadd_rn(Tn[c], t);
                            Kn[c] += cd.code;
                        }
                    } else if constexpr (kMode == 1) {
                        constexpr int x = BG::slot[e];
                        Lx[x] = __fadd_rn(Lx[x], cn[k]);
                    } else if constexpr (kMode == 2) {
                        constexpr int x = BG::slot[e];
                        xneg_it |= __fadd_rn(Lx[x], cn[k]) < 0.0f ? (1ull << x) : 0ull;
                    }
                });
            });
            static_for<0, NC>([&](auto kc) {
                Tc[decltype(kc)::value] = Tn[decltype(kc)::value];
                Kc[decltype(kc)::value] = Kn[decltype(kc)::value];
            });
#if LDPC_FAST_TMEM
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
#if LDPC_FAST_TMEM == 2
            if constexpr (kMode != 1) issue_group(IC<0>{});  // first row group of the next iteration
#endif
        };
        // parity checks of the current posteriors: core columns from Tc (BP: resolved), degree-1 columns from `xneg`
        auto all_checks_ok = [&](unsigned long long xneg) -> bool {
            unsigned hb[NC];
            static_for<0, NC>([&](auto kc) {
                constexpr int c = decltype(kc)::value;
                if constexpr (kAlgo == LDPC_ALGO_BP && kEarly) hb[c] = bp_resolve(Tc[c], Kc[c]) < 0.0f ? 1u : 0u;
                else hb[c] = Tc[c] < 0.0f ? 1u : 0u;
            });
            unsigned bad = 0;
            static_for<0, BG::kRows>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                unsigned par = 0;
                static_for<BG::row_ptr[i], BG::row_ptr[i + 1]>([&](auto ec) {
                    constexpr int e = decltype(ec)::value;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int s = BG::shift[e], c = BG::slot[e];
                        par ^= (s == 0) ? hb[c] : __shfl_sync(kFull, hb[c], lp[s], Z);
                    } else {
                        constexpr int x = BG::slot[e];
                        par ^= (unsigned)(xneg >> x) & 1u;
                    }
                });
                bad |= par;
            });
            const unsigned m = __ballot_sync(kFull, bad != 0);
            const unsigned gmask = (Z == 32) ? kFull : (((1u << (Z & 31)) - 1u) << (cwi * Z));
            return (m & gmask) == 0;
        };
        int iters_done = p.iters;
        bool early_ok = false;
        unsigned tneg_fz = 0;                                  // kEarly: decisions of the core columns, frozen at the codeword's stop
        if constexpr (kEarly) {
            bool frozen = false;
            unsigned long long xneg_fz = 0;
            for (int it = 1;; ++it) {
                xneg_it = 0;
                iteration(IC<2>{});
                const bool ok = all_checks_ok(xneg_it);        // of THIS lane's codeword
                if (!frozen && (ok || it >= p.iters)) {
                    frozen = true;
                    early_ok = ok;
                    iters_done = it;
                    xneg_fz = xneg_it;
                    static_for<0, NC>([&](auto kc) {
                        constexpr int c = decltype(kc)::value;
                        const float t = (kAlgo == LDPC_ALGO_BP) ? bp_resolve(Tc[c], Kc[c]) : Tc[c];
                        tneg_fz |= t < 0.0f ? (1u << c) : 0u;
                    });
                }
                if (__all_sync(kFull, frozen)) break;          // Z = 32: the one codeword of the warp
            }
            xneg_it = xneg_fz;
#if LDPC_FAST_TMEM == 2
            settle_group(IC<0>{});                             // the loads issued for an iteration that does not run
#endif
        } else {
            for (int it = 1; it < p.iters; ++it) iteration(IC<0>{});
            iteration(IC<1>{});
        }
        if constexpr (kAlgo == LDPC_ALGO_BP)
            static_for<0, NC>([&](auto kc) { Tc[decltype(kc)::value] = bp_resolve(Tc[decltype(kc)::value], Kc[decltype(kc)::value]); });

        // ---- outputs: Tc = posteriors of the core columns, Lx = posteriors of the degree-1 columns ----
        // (the output mode is tested once, not per column: each mode is its own unrolled store loop)
        auto belief_of = [&](auto jc) -> float {
            constexpr int j = decltype(jc)::value, sl = BG::col_slot[j];
            if constexpr (BG::col_kind[j] == 0) return Tc[sl];
            else return Lx[sl];
        };
        auto neg_of = [&](auto jc) -> bool {                  // hard decision of column block j at this lane
            constexpr int j = decltype(jc)::value, sl = BG::col_slot[j];
            if constexpr (kEarly && BG::col_kind[j] == 0) return ((tneg_fz >> sl) & 1u) != 0;
            else if constexpr (kEarly) return ((xneg_it >> sl) & 1ull) != 0;
            else if constexpr (BG::col_kind[j] == 0) return Tc[sl] < 0.0f;
            else return Lx[sl] < 0.0f;
        };
        if (!kEarly && p.soft_out && live) {
            float* o = p.soft_out + cw * N + r;
            static_for<0, BG::kCols>([&](auto jc) { o[decltype(jc)::value * Z] = belief_of(jc); });
        }
        if (p.hard_out) {
            if (p.hard_dtype == LDPC_HARD_PACKED) {
                unsigned hw[NWR];
#pragma unroll
                for (int q = 0; q < NWR; ++q) hw[q] = 0;
                static_for<0, BG::kCols>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    const unsigned b = __ballot_sync(kFull, neg_of(jc));
                    constexpr int wj = (j * Z) >> 5, off = (j * Z) & 31;
                    const unsigned mine = (Z == 32) ? b : ((b >> (cwi * Z)) & ((1u << (Z & 31)) - 1u));
                    if (r == wj % Z) hw[wj / Z] |= mine << off;
                });
                if (live) {
#pragma unroll
                    for (int q = 0; q < NWR; ++q)
                        if (q * Z + r < NW) ((unsigned*)p.hard_out)[cw * NW + q * Z + r] = hw[q];
                }
            } else if (p.hard_dtype == LDPC_HARD_F32) {
                if (live) {
                    float* o = (float*)p.hard_out + cw * N + r;
                    static_for<0, BG::kCols>([&](auto jc) { o[decltype(jc)::value * Z] = neg_of(jc) ? 1.0f : 0.0f; });
                }
            } else {
                if (live) {
                    uint8_t* o = (uint8_t*)p.hard_out + cw * N + r;
                    static_for<0, BG::kCols>([&](auto jc) { o[decltype(jc)::value * Z] = neg_of(jc) ? 1 : 0; });
                }
            }
        }
        if (p.iters_out && live && r == 0) p.iters_out[cw] = iters_done;
        if (p.syndrome_ok || p.counters) {
            unsigned long long xneg = 0;      // hard decisions of the degree-1 columns
            bool ok;
            if constexpr (kEarly) {
                xneg = xneg_it;
                ok = early_ok;
            } else {
                static_for<0, NX>([&](auto xc) { xneg |= Lx[decltype(xc)::value] < 0.0f ? (1ull << decltype(xc)::value) : 0ull; });
                ok = all_checks_ok(xneg);
            }
            if (p.syndrome_ok && live && r == 0) p.syndrome_ok[cw] = ok ? 1 : 0;
            if (p.counters) {
                // all-zero codeword was sent: every negative posterior is a bit error
                unsigned e = __popcll(xneg);
                if constexpr (kEarly) e += __popc(tneg_fz);
                else static_for<0, NC>([&](auto kc) { e += Tc[decltype(kc)::value] < 0.0f ? 1u : 0u; });
#pragma unroll
                for (int o = Z / 2; o > 0; o >>= 1) e += __shfl_xor_sync(kFull, e, o, Z);
                if (live && r == 0) {
                    acc_bits += e;
                    acc_fe += e != 0;
                    acc_frames += 1;
                    acc_und += (e != 0 && ok);
                }
            }
        }
    }
    if (p.counters) flush_counters(p.counters, acc_bits, acc_fe, acc_frames, acc_und);
#if LDPC_FAST_TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
#endif
}

}  // namespace ldpc
