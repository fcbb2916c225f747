// neural_qc_kernel.cuh -- the QC-structured LDPCNeuralDecoder kernel (design: neural_qc.cuh), rolled.
//
// A first version unrolled the whole base graph as decode_fast_kernel does.  It was bit-identical to the table-driven
// kernel and no faster: ncu (profiles/r2_ncu_neural_qc.md) showed 39 k warp-instructions per codeword at 12.7 % issue
// utilisation with `no_instruction` 3.9 stalls per issue -- 550 KB of straight-line SASS streamed from L2 by four
// warps.  This version keeps the HOT code inside the instruction cache and hides latency with more warps:
//   * four warps ("members") share one codeword: they sit in the same TMEM lane quarter (warp % 4), so all of them
//     reach the codeword's Tensor-Memory state; every base row (phase A) and base column (phase B) is owned by one
//     member (static, load-balanced schedule: csrc/nq_tables.h), members meet at a 128-thread named barrier between
//     phases.  16 warps per SM instead of 4;
//   * rows are processed by ONE unrolled body per row shape (2/3/4/5 core edges + a degree-1 edge, 8 or 10 core edges),
//     columns by one body per size class (6, 8, 10, 13, 16, 23 cells; shorter columns are zero-padded at the END of
//     their list, which leaves every fp32 sum unchanged); cell index and shift of an edge come from constant memory.
#pragma once
#include <math_constants.h>

#include "bg2_tables.h"
#include "neural_qc.cuh"
#include "nq_tables.h"
#include "params.cuh"

namespace ldpc {

constexpr int kNqGroups = 4;                      // codewords resident per CTA = TMEM lane quarters
constexpr int kNqThreads = kNqGroups * nq::kMembers * 32;
constexpr int kNqPitch = 33;                      // transposing tile: conflict-free [cell][lane] reads, <= 2-way conflicts on the row side
constexpr int kNqRowNc[nq::kRowClasses] = {2, 3, 4, 5, 8, 10};
constexpr int kNqRowNe[nq::kRowClasses] = {1, 1, 1, 1, 0, 0};
constexpr int kNqColMax[nq::kColClasses] = {6, 8, 10, 13, 16, 23};
constexpr int kNqColMin[nq::kColClasses] = {5, 7, 9, 12, 14, 22};   // smallest degree that uses the class (no masking below it)

template <class BG>
constexpr int nq_group_floats() { return BG::kEdges * kNqPitch + 3 * BG::kExtCols * 32 + 8; }
template <class BG>
constexpr size_t neural_qc_smem_bytes() {      // w_ch tile | per-codeword state | tile addresses of the row chunks (uint16)
    return sizeof(float) * ((size_t)BG::kEdges * 32 + (size_t)kNqGroups * nq_group_floats<BG>()) + sizeof(unsigned short) * BG::kEdges * 32;
}

__device__ __forceinline__ void nq_ld1_issue(uint32_t taddr, float& a) {
    uint32_t x;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr));
    a = u2f(x);
}
__device__ __forceinline__ void nq_tie(float& a) { asm volatile("" : "+f"(a)); }
__device__ __forceinline__ void nq_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(f2u(v)) : "memory");
}
__device__ __forceinline__ void nq_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void nq_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the members of a codeword meet: TMEM stores settled, ordered before / after the barrier, shared memory too
__device__ __forceinline__ void nq_group_sync(int group) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(nq::kMembers * 32) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <class BG>
__global__ void __launch_bounds__(kNqThreads, 1) neural_qc_kernel(const NeuralQcParams p) {
    static_assert(BG::kZ == 32, "one codeword per warp-wide lane set");
    constexpr int EB = BG::kEdges, EC = BG::kCoreEdges, NX = BG::kExtCols;
    constexpr int ECP = (EC + 3) / 4 * 4;
    constexpr int E = EB * 32;
    static_assert(EB == nq::kCells && 3 * ECP <= 512, "schedule tables / TMEM budget");
    extern __shared__ float nq_smem[];
    // the warp index is broadcast from lane 0 so that the compiler KNOWS it is warp-uniform: schedule-table indices,
    // TMEM addresses and loop bounds then live in uniform registers (the first build spent 5.5 k R2UR per codeword
    // moving per-access TMEM addresses into them)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int grp = warp & (kNqGroups - 1), mem = warp / kNqGroups;       // TMEM lane quarter = codeword slot; member in it
    float* wsm = nq_smem;                                                  // [EB][32]  w_ch per (cell, lane)
    float* lls = nq_smem + EB * 32 + grp * nq_group_floats<BG>();         // [EB][33]  llr_e, later the soft outputs
    float* xe0 = lls + EB * kNqPitch;                                      // [2][NX][32] ring of the degree-1 cells
    float* ces = xe0 + 2 * NX * 32;                                        // [NX][32]  their last check message
    float* red = ces + NX * 32;                                            // [4] per-member loss maxima
    // where element `lane` of the m-th 32-edge chunk of a row sits in the tile: (D + k) * 33 + r (the same for every codeword)
    unsigned short* tile_addr = reinterpret_cast<unsigned short*>(nq_smem + EB * 32 + kNqGroups * nq_group_floats<BG>());

    __shared__ uint32_t tmem_base_s;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // w_ch[e] -> wsm[cell][lane], e = 32*D_j + lane*d_j + k
    for (int m = warp; m < EB; m += kNqThreads / 32) {
        const unsigned cmeta = nq::chunk_meta[m];
        const int D = cmeta & 0xff, d = (cmeta >> 8) & 0x1f;
        wsm[m * 32 + lane] = __ldg(p.w_ch + 32 * D + lane * d + (m - D));
        const int inv = cmeta >> 13, off = 32 * (m - D) + lane, r = (off * inv) >> 16, k = off - r * d;
        tile_addr[m * 32 + lane] = (unsigned short)((D + k) * kNqPitch + r);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = __shfl_sync(kFull, tmem_base_s, 0) + (((uint32_t)grp * 32u) << 16);
    const uint32_t tC = tbase;                          // c2v
    uint32_t tXc = tbase + ECP, tXo = tbase + 2 * ECP;  // ring: current x, older x (roles swap every iteration)
    int xc_off = 0, xo_off = NX * 32;                   // the same for the degree-1 ring in shared memory
    const float wres0 = p.L >= 1 ? __ldg(p.w_res) : 0.0f, wres1 = p.L >= 2 ? __ldg(p.w_res + 1) : 0.0f;

    for (long long cw0 = (long long)blockIdx.x * kNqGroups; cw0 < p.B; cw0 += (long long)gridDim.x * kNqGroups) {
        // groups past the end of the batch keep walking (on the last codeword, without storing): every barrier is reached
        const bool live = cw0 + grp < p.B;
        const long long cw = live ? cw0 + grp : p.B - 1;
        // ---- load llr_e[cw]: coalesced 128-byte chunks -> [cell][lane] tile ----
        {
            const float* src = p.llr + cw * E + lane;
#pragma unroll 8
            for (int m = mem; m < EB; m += nq::kMembers) lls[tile_addr[m * 32 + lane]] = __ldg(src + 32 * m);
        }
        nq_group_sync(grp);
        // x_0 = llr_e in the current ring slot, zeros in the older one (its residual weight is zero until it is written)
        for (int m = mem; m < EC; m += nq::kMembers) {
            nq_st1(tXc + m, lls[m * kNqPitch + lane]);
            nq_st1(tXo + m, 0.0f);
        }
        for (int x = mem; x < NX; x += nq::kMembers) {
            xe0[xc_off + x * 32 + lane] = lls[(EC + x) * kNqPitch + lane];
            xe0[xo_off + x * 32 + lane] = 0.0f;
        }
        nq_group_sync(grp);

        // ---- phase A: CheckLayer on the current x; one body per row shape ----
        auto row_body = [&](auto ncc, auto nec, int row, bool last) {
            constexpr int NC = decltype(ncc)::value, NE = decltype(nec)::value, d = NC + NE;
            float v[d], a[d];
            unsigned sb[d];
            int zi[d], cell[NC], sft[NC];
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const unsigned meta = nq::row_meta[row][k];
                cell[k] = meta & 0xff;
                sft[k] = meta >> 8;
                nq_ld1_issue(tXc + cell[k], v[k]);
            });
            nq_wait_ld();
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                nq_tie(v[k]);
                v[k] = __shfl_sync(kFull, v[k], lane + sft[k]);           // variable (r + s) mod 32 -> check row r
            });
            int xs = 0;
            if constexpr (NE) {
                xs = nq::row_ext[row];
                v[NC] = xe0[xc_off + xs * 32 + lane];
            }
            unsigned nb = 0;
            int zc = 0;
            float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                // neural_check_visit (neural.cuh): sign(v + 1e-10) factors as (xor of sign bits, "a factor was 0"),
                // zeros count as 1e10 in the minimum
                const float sh = __fadd_rn(v[k], 1e-10f);
                sb[k] = f2u(sh);
                nb ^= sb[k];
                zi[k] = !(fabsf(sh) > 0.0f) ? 1 : 0;
                zc += zi[k];
                const float av = fabsf(v[k]);
                a[k] = av > 0.0f ? av : 1e10f;
                m2 = fminf(m2, fmaxf(m1, a[k]));
                m1 = fminf(m1, a[k]);
            });
            // fewer than 9 other edges: the table's padded slots are zero inputs, magnitude 1e10 (layers.py:48-57)
            const float m1c = (d - 1 < 9) ? fminf(m1, 1e10f) : m1, m2c = (d - 1 < 9) ? fminf(m2, 1e10f) : m2;
            // a zero FACTOR needs v = -1e-10f exactly (or NaN): practically never.  Without one the sign product is +-1 and
            // the message is the minimum with the sign bits xor-ed in (what the multiplication by +-1.0f yields); the
            // general form runs only when some lane of the warp saw a zero factor or a non-finite minimum.
            const bool special = __any_sync(kFull, zc != 0 || !(m2c < CUDART_INF_F));
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                if (k < NC || last) {
                    const float m = (a[k] == m1) ? m2c : m1c;                       // minimum over the OTHER edges
                    const unsigned sgn = (nb ^ sb[k]) & 0x80000000u;               // their sign product
                    float o = u2f(sgn | f2u(m));
                    if (special) o = __fmul_rn(u2f(sgn | ((zc - zi[k]) > 0 ? 0u : 0x3f800000u)), m);
                    if constexpr (k < NC) nq_st1(tC + cell[k], __shfl_sync(kFull, o, lane - sft[k]));
                    else ces[xs * 32 + lane] = o;
                }
            });
        };
        auto phase_a = [&](bool last) {
            static_for<0, nq::kRowClasses>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const int t1 = nq::sched_row_ptr[mem][c + 1];
#pragma unroll 1
                for (int t = nq::sched_row_ptr[mem][c]; t < t1; ++t)
                    row_body(IC<kNqRowNc[c]>{}, IC<kNqRowNe[c]>{}, (int)nq::sched_rows[mem][t], last);
            });
        };

        // ---- phase B / final: one base column, lane-local.  Sums over the OTHER edges of a variable in the table's
        //      (ascending check) order, starting from 0 (neural_gather_sum): running prefix + suffix chain ----
        auto col_body = [&](auto dmc, auto dnc, auto finalc, int b0, int d, float wr0, float wr1) {
            constexpr int DM = decltype(dmc)::value, DMIN = decltype(dnc)::value;
            constexpr bool kFinal = decltype(finalc)::value != 0;
            float c[DM], xc[DM], xo[DM];
            float* lcol = lls + b0 * kNqPitch + lane;          // this lane's LLRs / soft outputs of the column, pitch 33
            const float* wcol = wsm + b0 * 32 + lane;
            const uint32_t tcol = tXo + b0;
            {
                const uint32_t ta = tC + b0, tb = tXc + b0;
                static_for<0, DM>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    nq_ld1_issue(ta + k, c[k]);
                    if constexpr (!kFinal) {
                        nq_ld1_issue(tb + k, xc[k]);
                        nq_ld1_issue(tcol + k, xo[k]);
                    }
                });
            }
            nq_wait_ld();
            static_for<0, DM>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                nq_tie(c[k]);
                if constexpr (!kFinal) { nq_tie(xc[k]); nq_tie(xo[k]); }
                if constexpr (k >= DMIN) c[k] = k < d ? c[k] : 0.0f;     // cells of the NEXT column: zero, at the end of the list
            });
            float pre = 0.0f;
            static_for<0, DM>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                float s = pre;
                static_for<k + 1, DM>([&](auto k2) { s = __fadd_rn(s, c[decltype(k2)::value]); });
                pre = __fadd_rn(pre, c[k]);
                if (k < DMIN || k < d) {
                    const float ll = lcol[k * kNqPitch];
                    if constexpr (kFinal) {
                        const float z = __fadd_rn(__fadd_rn(c[k], s), ll);
                        lcol[k * kNqPitch] = 1.0f / (1.0f + expf(-z));
                    } else {
                        float r = __fadd_rn(__fmul_rn(ll, wcol[k * 32]), s);
                        r = __fadd_rn(r, __fmul_rn(wr0, xc[k]));
                        r = __fadd_rn(r, __fmul_rn(wr1, xo[k]));
                        nq_st1(tcol + k, r);
                    }
                }
            });
        };
        auto phase_cols = [&](auto finalc, float wr0, float wr1) {
            static_for<0, nq::kColClasses>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const int t1 = nq::sched_col_ptr[mem][c + 1];
#pragma unroll 1
                for (int t = nq::sched_col_ptr[mem][c]; t < t1; ++t) {
                    const int j = nq::sched_cols[mem][t];
                    col_body(IC<kNqColMax[c]>{}, IC<kNqColMin[c]>{}, finalc, (int)nq::col_b0[j], (int)nq::col_d[j], wr0, wr1);
                }
            });
            const int nx = nq::sched_ext_cnt[mem];
#pragma unroll 1
            for (int t = 0; t < nx; ++t) {
                const int x = nq::sched_ext[mem][t], m = EC + x;
                const float ll = lls[m * kNqPitch + lane];
                if constexpr (decltype(finalc)::value != 0) {
                    const float z = __fadd_rn(__fadd_rn(ces[x * 32 + lane], 0.0f), ll);
                    lls[m * kNqPitch + lane] = 1.0f / (1.0f + expf(-z));
                } else {
                    float r = __fadd_rn(__fmul_rn(ll, wsm[m * 32 + lane]), 0.0f);          // a degree-1 variable has no other edge
                    r = __fadd_rn(r, __fmul_rn(wr0, xe0[xc_off + x * 32 + lane]));
                    r = __fadd_rn(r, __fmul_rn(wr1, xe0[xo_off + x * 32 + lane]));
                    xe0[xo_off + x * 32 + lane] = r;
                }
            }
        };

        for (int l = 0; l < p.iters; ++l) {
            if (p.save_x && live) {
                // training forward: the input of this CheckLayer in the caller's edge order (d_j-strided 4-byte stores,
                // merged by the L2: each column's block of 32*d_j floats is written completely by one member)
                float* dst = p.save_x + ((long long)l * p.B + cw) * E;
                for (int m = mem; m < EC; m += nq::kMembers) {
                    const unsigned cmeta = nq::chunk_meta[m];
                    const int D = cmeta & 0xff, d = (cmeta >> 8) & 0x1f;
                    float xv;
                    nq_ld1_issue(tXc + m, xv);
                    nq_wait_ld();
                    nq_tie(xv);
                    dst[32 * D + lane * d + (m - D)] = xv;
                }
                for (int x = mem; x < NX; x += nq::kMembers) dst[32 * (EC + x) + lane] = xe0[xc_off + x * 32 + lane];
            }
            const bool last = l == p.iters - 1;
            phase_a(last);
            nq_group_sync(grp);
            if (last) break;
            // queue of earlier outputs (models/decoder.py): x_0 is not an entry, so the first update has no residual term
            phase_cols(IC<0>{}, l >= 1 ? wres0 : 0.0f, l >= 2 ? wres1 : 0.0f);
            nq_group_sync(grp);
            { const uint32_t t = tXc; tXc = tXo; tXo = t; }
            { const int t = xc_off; xc_off = xo_off; xo_off = t; }
        }
        // ---- final = VariableLayer(c2v, c2v); OutputLayer: soft = sigmoid(final + llr), staged in the tile ----
        phase_cols(IC<1>{}, 0.0f, 0.0f);
        nq_group_sync(grp);
        {
            float* dst = p.soft + cw * E + lane;
            const float* gts = p.gt ? p.gt + cw * E + lane : nullptr;
            float best = -CUDART_INF_F;
#pragma unroll 2
            for (int m = mem; m < EB; m += nq::kMembers) {
                const float s = lls[tile_addr[m * 32 + lane]];
                if (live) dst[32 * m] = s;
                if (gts) {
                    const float y = __ldg(gts + 32 * m);
                    const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                    const float loss = -(y * l1 + (1.0f - y) * l0);
                    best = loss > best ? loss : best;
                }
            }
            if (gts) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(kFull, best, o));
                if (lane == 0) red[mem] = best;
            }
        }
        nq_group_sync(grp);                                  // tile and `red` are complete; the next codeword may overwrite the tile
        if (p.gt && live && mem == 0 && lane == 0)
            p.max_loss[cw] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        nq_group_sync(grp);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

}  // namespace ldpc
