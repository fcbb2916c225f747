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
//     iterations as per-lane private columns msg[e][lane] (conflict-free, no barriers);
//   * check node: running min1/min2 + XOR of sign bits; the message to edge k is
//     alpha*(|v_k|==min1 ? min2 : min1) with sign (total ^ sign_k) -- identical to the
//     reference's product-of-signs / min-over-others including its sign(0)=0 rule, because a
//     zero input is the minimum and zeroes every other output by magnitude;
//   * variable node: v2c = T - c2v (total minus self) instead of the reference's sum over the
//     other checks: this is the one place the operation order differs (<= a few ulp of T per
//     iteration; the posterior itself is accumulated in the reference's ascending-check
//     order).  LDPC_PATH_EXACT keeps the reference order bit for bit.
#pragma once
#include <math_constants.h>

#include "bg2_tables.h"
#include "decode_exact.cuh"

namespace ldpc {

template <int I>
struct IC {
    static constexpr int value = I;
};
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr unsigned kFull = 0xffffffffu;

template <class BG, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) minsum_fast_kernel(const DecodeParams p) {
    constexpr int Z = BG::kZ, G = 32 / Z, NC = BG::kCoreCols, NX = BG::kExtCols, EC = BG::kCoreEdges;
    constexpr int N = BG::kCols * Z, NW = (N + 31) / 32, NWR = (NW + Z - 1) / Z;
    static_assert(NX <= 64, "degree-1 column bitmap is 64 bits");
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* msg = smem + warp * ((EC + NC) * 32) + lane;   // msg[e*32]: this lane's private column
    float* Ls = msg + EC * 32;                            // channel LLR of the core columns
    const int cwi = lane / Z, r = lane % Z;
    const float alpha = p.alpha;
    unsigned long long acc_bits = 0, acc_fe = 0, acc_frames = 0, acc_und = 0;

    for (long long grp = (long long)blockIdx.x * kWarps + warp; grp < p.ngroups; grp += (long long)gridDim.x * kWarps) {
        const long long cw = grp * G + cwi;
        const bool live = cw < p.B;
        float Tc[NC], Lx[NX];
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
                        float z[4];
                        normal4(p.gen.seed, frame, (uint32_t)(hi * 32 + jm * Z + r), z);
                        static_for<0, 4>([&](auto cc) {
                            constexpr int j = (hi * 4 + decltype(cc)::value) * P + jm;
                            if constexpr (j < BG::kCols)
                                put_llr(IC<j>{}, live ? llr_from_noise(z[decltype(cc)::value], 1.0f, p.gen) : 0.0f);
                        });
                    }
                });
            });
        } else {
            const float* llr = p.llr + cw * N + r;
            static_for<0, BG::kCols>([&](auto jc) {
                put_llr(jc, live ? __ldg(llr + decltype(jc)::value * Z) : 0.0f);
            });
        }
        static_for<0, EC>([&](auto ec) { msg[decltype(ec)::value * 32] = 0.0f; });

        unsigned long long xneg = 0;   // hard decisions of the degree-1 columns (last iteration)
        unsigned hw[NWR];              // packed hard-decision words held by this lane
#pragma unroll
        for (int q = 0; q < NWR; ++q) hw[q] = 0;
        float* soft = p.soft_out ? p.soft_out + cw * N + r : nullptr;

        auto put_hard = [&](auto jc, float belief) {
            constexpr int j = decltype(jc)::value;
            const bool neg = belief < 0.0f;
            if (p.hard_out) {
                if (p.hard_dtype == LDPC_HARD_F32) {
                    if (live) ((float*)p.hard_out)[cw * N + j * Z + r] = neg ? 1.0f : 0.0f;
                } else if (p.hard_dtype == LDPC_HARD_U8) {
                    if (live) ((uint8_t*)p.hard_out)[cw * N + j * Z + r] = neg ? 1 : 0;
                } else {
                    const unsigned b = __ballot_sync(kFull, neg);
                    constexpr int wj = (j * Z) >> 5, off = (j * Z) & 31;
                    const unsigned mine = (Z == 32) ? b : ((b >> (cwi * Z)) & ((1u << (Z & 31)) - 1u));
                    if (r == wj % Z) hw[wj / Z] |= mine << off;
                }
            }
        };

        for (int it = 0; it < p.iters; ++it) {
            const bool last = it == p.iters - 1;
            float Tn[NC];
            static_for<0, NC>([&](auto kc) { Tn[decltype(kc)::value] = Ls[decltype(kc)::value * 32]; });

            static_for<0, BG::kRows>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int e0 = BG::row_ptr[i], d = BG::row_ptr[i + 1] - e0;
                float v[d];
                // gather variable-to-check messages, check-aligned
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int c = BG::slot[e], s = BG::shift[e], mi = BG::msg[e];
                        const float t = (s == 0) ? Tc[c] : __shfl_sync(kFull, Tc[c], lane + s, Z);
                        v[k] = t - msg[mi * 32];
                    } else {
                        static_assert(BG::kind[e] == 0 || BG::shift[e] == 0, "degree-1 columns are expected unshifted");
                        constexpr int x = BG::slot[e];
                        v[k] = Lx[x];
                    }
                });
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
                sg &= 0x80000000u;
                const unsigned p1 = f2u(__fmul_rn(alpha, m1)) ^ sg, p2 = f2u(__fmul_rn(alpha, m2)) ^ sg;
                // emit check-to-variable messages, accumulate the new posteriors
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int c = BG::slot[e], s = BG::shift[e], mi = BG::msg[e];
                        const unsigned sel = (fabsf(v[k]) == m1) ? p2 : p1;
                        const float cnew = u2f(sel ^ (f2u(v[k]) & 0x80000000u));
                        msg[mi * 32] = cnew;
                        const float t = (s == 0) ? cnew : __shfl_sync(kFull, cnew, lane + (Z - s), Z);
                        Tn[c] = __fadd_rn(Tn[c], t);
                    } else {
                        if (last) {
                            constexpr int x = BG::slot[e], jcol = BG::col[e];
                            const unsigned sel = (fabsf(v[k]) == m1) ? p2 : p1;
                            const float belief = __fadd_rn(Lx[x], u2f(sel ^ (f2u(v[k]) & 0x80000000u)));
                            if (soft && live) soft[jcol * Z] = belief;
                            if (belief < 0.0f) xneg |= 1ull << x;
                            put_hard(IC<jcol>{}, belief);
                        }
                    }
                });
            });
            static_for<0, NC>([&](auto kc) { Tc[decltype(kc)::value] = Tn[decltype(kc)::value]; });
        }

        // ---- outputs of the core columns ----
        static_for<0, NC>([&](auto kc) {
            constexpr int k = decltype(kc)::value, j = BG::core_col[k];
            if (soft && live) soft[j * Z] = Tc[k];
            put_hard(IC<j>{}, Tc[k]);
        });
        if (p.hard_out && p.hard_dtype == LDPC_HARD_PACKED && live) {
#pragma unroll
            for (int q = 0; q < NWR; ++q)
                if (q * Z + r < NW) ((unsigned*)p.hard_out)[cw * NW + q * Z + r] = hw[q];
        }
        if (p.iters_out && live && r == 0) p.iters_out[cw] = p.iters;
        if (p.syndrome_ok || p.counters) {
            unsigned bad = 0;
            static_for<0, BG::kRows>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                unsigned par = 0;
                static_for<BG::row_ptr[i], BG::row_ptr[i + 1]>([&](auto ec) {
                    constexpr int e = decltype(ec)::value;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int s = BG::shift[e], c = BG::slot[e];
                        const unsigned b = Tc[c] < 0.0f ? 1u : 0u;
                        par ^= (s == 0) ? b : __shfl_sync(kFull, b, lane + s, Z);
                    } else {
                        constexpr int x = BG::slot[e];
                        par ^= (unsigned)(xneg >> x) & 1u;
                    }
                });
                bad |= par;
            });
            const unsigned m = __ballot_sync(kFull, bad != 0);
            const unsigned gmask = (Z == 32) ? kFull : (((1u << (Z & 31)) - 1u) << (cwi * Z));
            const bool ok = (m & gmask) == 0;
            if (p.syndrome_ok && live && r == 0) p.syndrome_ok[cw] = ok ? 1 : 0;
            if (p.counters) {
                // all-zero codeword was sent: every negative posterior is a bit error
                unsigned e = __popcll(xneg);
                static_for<0, NC>([&](auto kc) { e += Tc[decltype(kc)::value] < 0.0f ? 1u : 0u; });
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
}

// ---- host side ----------------------------------------------------------------------------
template <class BG>
inline bool table_matches(const ldpc_code* c) {
    if (c->rows != BG::kRows || c->cols != BG::kCols || c->Z != BG::kZ || c->E != BG::kEdges) return false;
    const uint32_t* t = c->h_tab.data();
    const int off_rowptr = t[7], off_redge = t[9];
    for (int i = 0; i <= BG::kRows; ++i)
        if ((int)t[off_rowptr + i] != BG::row_ptr[i]) return false;
    for (int e = 0; e < BG::kEdges; ++e) {
        const uint32_t w = t[off_redge + e];
        if ((int)(w & 0xffff) != BG::col[e] || (int)((w >> 16) & 0xff) != BG::shift[e]) return false;
    }
    return true;
}

inline int detect_fast_kind(const ldpc_code* c) {
    if (table_matches<BG2Z32>(c)) return 1;
    if (table_matches<BG2Z4>(c)) return 2;
    return 0;
}

inline bool fast_path_supports(const ldpc_code* c, int algo, int stop_mode, bool want_mask) {
    return c->fast_kind != 0 && algo == LDPC_ALGO_MINSUM && stop_mode == LDPC_STOP_FIXED && !want_mask;
}

constexpr int kFastWarps = 10;

template <class BG>
inline int launch_fast_inst(DecodeParams p, cudaStream_t st) {
    constexpr int G = 32 / BG::kZ;
    constexpr size_t smem = (size_t)kFastWarps * (BG::kCoreEdges + BG::kCoreCols) * 32 * sizeof(float);
    static_assert(smem <= (size_t)kMaxSmemPerBlock, "fast kernel shared memory");
    auto kern = minsum_fast_kernel<BG, kFastWarps>;
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.ngroups = (p.B + G - 1) / G;
    long long blocks = (p.ngroups + kFastWarps - 1) / kFastWarps;
    if (blocks > kNumSMs) blocks = kNumSMs;
    kern<<<(int)blocks, kFastWarps * 32, smem, st>>>(p);
    LDPC_CHECK_LAUNCH("minsum_fast_kernel");
    return LDPC_OK;
}

inline int launch_fast(const ldpc_code* c, int algo, const DecodeParams& p, cudaStream_t st) {
    (void)algo;
    if (c->fast_kind == 1) return launch_fast_inst<BG2Z32>(p, st);
    if (c->fast_kind == 2) return launch_fast_inst<BG2Z4>(p, st);
    return fail(LDPC_ERR_UNSUPPORTED, "fast path: code is not one of the compiled tables");
}

}  // namespace ldpc
