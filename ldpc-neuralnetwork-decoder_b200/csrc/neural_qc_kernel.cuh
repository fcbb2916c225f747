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
constexpr int nq_group_floats() { return BG::kEdges * kNqPitch + 3 * BG::kExtCols * 32 + 128; }
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
// the members of a codeword meet: TMEM stores settled, ordered before / after the barrier, shared memory too
__device__ __forceinline__ void nq_group_sync(int group) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(nq::kMembers * 32) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <class BG>
__global__ void __launch_bounds__(kNqThreads, 1) neural_qc_kernel(const NeuralQcParams p) {
    // Z < 32: a warp holds G = 32 / Z codewords side by side (lane = Z * sub + r); every per-lane structure (Tensor Memory
    // columns, tile columns) then carries G codewords at once, rotations stay inside a codeword's Z lanes, and only the global
    // addressing, the rotations and the per-codeword reductions know about it.  Z = 32 compiles to the code it always was.
    constexpr int Z = BG::kZ, G = 32 / Z;
    static_assert(Z == 32 || Z == 16 || Z == 8 || Z == 4, "lifting sizes with a whole number of codewords per warp");
    constexpr int EB = BG::kEdges, EC = BG::kCoreEdges, NX = BG::kExtCols;
    constexpr int ECP = (EC + 3) / 4 * 4;
    constexpr int E = EB * Z, NV = BG::kCols * Z;
    static_assert(EB == nq::kCells && 3 * ECP <= 512, "schedule tables / TMEM budget");
    extern __shared__ float nq_smem[];
    // the warp index is broadcast from lane 0 so that the compiler KNOWS it is warp-uniform: schedule-table indices,
    // TMEM addresses and loop bounds then live in uniform registers (the first build spent 5.5 k R2UR per codeword
    // moving per-access TMEM addresses into them)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int grp = warp & (kNqGroups - 1), mem = warp / kNqGroups;       // TMEM lane quarter = codeword slot; member in it
    const int r = lane & (Z - 1), sub = lane / Z, lbase = lane & ~(Z - 1); // circulant row, codeword of the warp, its first lane
    // rotation by s inside the codeword's lanes: lane r reads row (r + s) mod Z
    auto rot = [&](float v, int s) {
        if constexpr (Z == 32) return __shfl_sync(kFull, v, lane + s);
        else return __shfl_sync(kFull, v, lbase | ((lane + s) & (Z - 1)));
    };
    auto rot_back = [&](float v, int s) {                                  // the inverse: lane r reads row (r - s) mod Z
        if constexpr (Z == 32) return __shfl_sync(kFull, v, lane - s);
        else return __shfl_sync(kFull, v, lbase | ((lane - s) & (Z - 1)));
    };
    float* wsm = nq_smem;                                                  // [EB][32]  w_ch per (cell, lane)
    float* lls = nq_smem + EB * 32 + grp * nq_group_floats<BG>();         // [EB][33]  llr_e, later the soft outputs
    float* xe0 = lls + EB * kNqPitch;                                      // [2][NX][32] ring of the degree-1 cells
    float* ces = xe0 + 2 * NX * 32;                                        // [NX][32]  their last check message
    float* red = ces + NX * 32;                                            // [G][4][4] per codeword: per-member loss maximum, its edge, soft and target
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
        wsm[m * 32 + lane] = __ldg(p.w_ch + Z * D + r * d + (m - D));
        const int inv = cmeta >> 13, off = Z * (m - D) + r, rr = (off * inv) >> 16, k = off - rr * d;
        tile_addr[m * 32 + lane] = (unsigned short)((D + k) * kNqPitch + lbase + rr);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = __shfl_sync(kFull, tmem_base_s, 0) + (((uint32_t)grp * 32u) << 16);
    const uint32_t tC = tbase;                          // c2v
    uint32_t tXc = tbase + ECP, tXo = tbase + 2 * ECP;  // ring: current x, older x (roles swap every iteration)
    int xc_off = 0, xo_off = NX * 32;                   // the same for the degree-1 ring in shared memory
    const float wres0 = p.L >= 1 ? __ldg(p.w_res) : 0.0f, wres1 = p.L >= 2 ? __ldg(p.w_res + 1) : 0.0f;

    for (long long cw0 = (long long)blockIdx.x * kNqGroups * G; cw0 < p.B; cw0 += (long long)gridDim.x * kNqGroups * G) {
        // groups past the end of the batch keep walking (on the last codeword, without storing): every barrier is reached
        const bool live = cw0 + grp * G + sub < p.B;
        const long long cw = live ? cw0 + grp * G + sub : p.B - 1;
        // the NEXT codeword of this group: pull its rows (llr_e, ground truth) into L2 while this one is decoded
        {
            const long long nxt = cw0 + ((long long)gridDim.x * kNqGroups + grp) * G;     // first of the group's next G codewords
            if (nxt < p.B) {
                const int t = mem * 32 + lane;                      // 128 threads, one 128-byte line each per step
                const long long rowlen = p.per_var ? NV : E;
                const long long have = (p.B - nxt < G ? p.B - nxt : G) * rowlen;       // the G rows are adjacent in memory
                for (long long i = 32 * t; i < have; i += 32 * nq::kMembers * 32) {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(p.llr + nxt * rowlen + i));
                    if (p.gt) asm volatile("prefetch.global.L2 [%0];" :: "l"(p.gt + nxt * rowlen + i));
                }
            }
        }
        // ---- load llr_e[cw]: coalesced 128-byte chunks -> [cell][lane] tile ----
        if (p.per_var) {
            // one line per base column, copied to each of the column's cells (the value llr[:, edge_to_var] gives its edges);
            // a member's 13 lines are requested together
            const float* src = p.llr + cw * NV + r;
            constexpr int kPer = BG::kCols / nq::kMembers;
            static_assert(kPer * nq::kMembers == BG::kCols, "columns split evenly over the members");
            float v[kPer];
#pragma unroll
            for (int i = 0; i < kPer; ++i) v[i] = __ldg(src + Z * (mem + i * nq::kMembers));
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int j = mem + i * nq::kMembers;
                const int D = j < BG::kCoreCols ? (int)nq::col_b0[j] : EC + (j - BG::kCoreCols);
                const int d = j < BG::kCoreCols ? (int)nq::col_d[j] : 1;
#pragma unroll 1
                for (int k = 0; k < d; ++k) lls[(D + k) * kNqPitch + lane] = v[i];
            }
        } else {
            const float* src = p.llr + cw * E + r;
#pragma unroll 8
            for (int m = mem; m < EB; m += nq::kMembers) lls[tile_addr[m * 32 + lane]] = __ldg(src + Z * m);
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
        auto row_body = [&](auto ncc, auto nec, int row, bool last, float* save) {
            constexpr int NC = decltype(ncc)::value, NE = decltype(nec)::value, d = NC + NE;
            float v[d], a[d];
            unsigned sb[d];
            int zi[d], cell[NC], sft[NC];
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const unsigned meta = nq::row_meta[row][k];
                cell[k] = meta & 0xff;
                sft[k] = Z == 32 ? (meta >> 8) : ((meta >> 8) & (Z - 1));
                nq_ld1_issue(tXc + cell[k], v[k]);
            });
            nq_wait_ld();
            static_for<0, NC>([&](auto kc) { nq_tie(v[decltype(kc)::value]); });
            // training forward: this CheckLayer's input as [cell][lane] (one coalesced 128-byte line per cell; every cell
            // belongs to exactly one row) -- the layout the backward kernel reads back without a transposing tile
            if (save) static_for<0, NC>([&](auto kc) { save[Z * cell[decltype(kc)::value]] = v[decltype(kc)::value]; });
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                v[k] = rot(v[k], sft[k]);                                  // variable (r + s) mod Z -> check row r
            });
            int xs = 0;
            if constexpr (NE) {
                xs = nq::row_ext[row];
                v[NC] = xe0[xc_off + xs * 32 + lane];
                if (save) save[Z * (EC + xs)] = v[NC];
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
                    if constexpr (k < NC) nq_st1(tC + cell[k], rot_back(o, sft[k]));
                    else ces[xs * 32 + lane] = o;
                }
            });
        };
        auto phase_a = [&](bool last, float* save) {
            static_for<0, nq::kRowClasses>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const int t1 = nq::sched_row_ptr[mem][c + 1];
#pragma unroll 1
                for (int t = nq::sched_row_ptr[mem][c]; t < t1; ++t)
                    row_body(IC<kNqRowNc[c]>{}, IC<kNqRowNe[c]>{}, (int)nq::sched_rows[mem][t], last, save);
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
            float* save = (p.save_x && live) ? p.save_x + ((long long)l * p.B + cw) * E + r : nullptr;      // [iteration][codeword][cell][Z]
            const bool last = l == p.iters - 1;
            phase_a(last, save);
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
        float best = -CUDART_INF_F, best_s = 0.0f, best_y = 0.0f;
        int besti = 0x7fffffff;
        if (p.per_var) {
            // cell by cell as below, in the tile's own [cell][lane] order: the output of a column's FIRST cell is the variables'
            // output line; every cell's loss is taken against the column's target line (L1 hits after the first cell);
            // edge (cell m of the column starting at D with degree d, variable `lane`) = 32 D + lane d + (m - D)
            float* dst = p.soft + cw * NV + r;
            const float* gts = p.gt ? p.gt + cw * NV + r : nullptr;
#pragma unroll 1
            for (int m0 = mem; m0 < EB; m0 += 4 * nq::kMembers) {
                float sv[4], yv[4];
                int col[4];
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const int m = m0 + i * nq::kMembers;
                    col[i] = m < EB ? (int)nq::cell_col[m] : 0;
                    sv[i] = m < EB ? lls[m * kNqPitch + lane] : 0.5f;
                    yv[i] = (gts && m < EB) ? __ldg(gts + Z * col[i]) : 0.0f;
                });
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const int m = m0 + i * nq::kMembers;
                    if (m < EB) {
                        const unsigned cmeta = nq::chunk_meta[m];
                        const int D = cmeta & 0xff, d = (cmeta >> 8) & 0x1f;
                        const float s = sv[i];
                        if (live && m == D) dst[Z * col[i]] = s;
                        if (gts) {
                            const float y = yv[i];
                            const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                            const float loss = -(y * l1 + (1.0f - y) * l0);
                            if (loss > best) { best = loss; besti = Z * D + r * d + (m - D); best_s = s; best_y = y; }   // ascending per thread
                        }
                    }
                });
            }
        } else {
            float* dst = p.soft + cw * E + r;
            const float* gts = p.gt ? p.gt + cw * E + r : nullptr;
            // four chunks at a time: their staged values and ground-truth lines are requested before the first logarithm
            // (training step 9.4 -> 8.5 ms per 32 768 codewords; a separate plain store loop for the no-ground-truth case
            // measured slower for both cases -- the kernel is sensitive to the code layout of its hot loops)
#pragma unroll 1
            for (int m0 = mem; m0 < EB; m0 += 4 * nq::kMembers) {
                float sv[4], yv[4];
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const int m = m0 + i * nq::kMembers;
                    sv[i] = m < EB ? lls[tile_addr[m * 32 + lane]] : 0.5f;
                    yv[i] = (gts && m < EB) ? __ldg(gts + Z * m) : 0.0f;
                });
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const int m = m0 + i * nq::kMembers;
                    if (m < EB) {
                        const float s = sv[i];
                        if (live) dst[Z * m] = s;
                        if (gts) {
                            // (a one-logarithm form for binary targets was measured SLOWER: 108 instead of 95 registers, inference
                            // 2.91 vs 2.60 ms per 32 768 codewords even though that path does not run without ground truth)
                            const float y = yv[i];
                            const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                            const float loss = -(y * l1 + (1.0f - y) * l0);
                            if (loss > best) { best = loss; besti = Z * m + r; best_s = s; best_y = y; }   // first (lowest) edge among equal maxima
                        }
                    }
                });
            }
        }
        if (p.gt) {
#pragma unroll
            for (int o = Z / 2; o > 0; o >>= 1) {                    // within the codeword's Z lanes
                const float ob = __shfl_xor_sync(kFull, best, o);
                const int oi = __shfl_xor_sync(kFull, besti, o);
                const float os = __shfl_xor_sync(kFull, best_s, o), oy = __shfl_xor_sync(kFull, best_y, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; best_s = os; best_y = oy; }
            }
            if (r == 0) { float* rd = red + 16 * sub; rd[mem] = best; rd[4 + mem] = __int_as_float(besti); rd[8 + mem] = best_s; rd[12 + mem] = best_y; }
        }
        nq_group_sync(grp);                                  // tile and `red` are complete; the next codeword may overwrite the tile
        if (p.gt && live && mem == 0 && r == 0) {
            const float* rd = red + 16 * sub;
            int q = 0;
#pragma unroll
            for (int t = 1; t < nq::kMembers; ++t) {
                const float ob = rd[t], cb = rd[q];
                const int oi = __float_as_int(rd[4 + t]), ci = __float_as_int(rd[4 + q]);
                if (ob > cb || (ob == cb && oi < ci)) q = t;
            }
            p.max_loss[cw] = rd[q];
            if (p.argmax) p.argmax[cw] = __float_as_int(rd[4 + q]);
            if (p.star) { p.star[2 * cw] = rd[8 + q]; p.star[2 * cw + 1] = rd[12 + q]; }
        }
        nq_group_sync(grp);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}


// ---- backward: d(sum_b g_ml[b] * max_loss[b]) / d(w_ch, w_res) from the saved CheckLayer inputs ---------------------------
// The autograd graph of the composition (models/decoder.py), walked per codeword with the same 4-warp / TMEM-quarter
// organisation as the forward:
//   OutputLayer: the frame's loss is the BCE of ONE edge e* (the arg-max): gz = g_ml (s-y)/max(s(1-s),1e-12) s(1-s)
//     (torch's own BCE and sigmoid backward formulas, as models/layers.py:_OutputFn uses them); final = sum of ALL check
//     messages of the variable, so every edge of e*'s variable receives gz as d/d c2v_last;
//   level l = I-1 .. 1:
//     CheckLayer backward (per base row): out_n = sign_n * min_{k != n} |x_k| sends its gradient to the first arg-min
//     k1 (from every n != k1) and to the runner-up k2 (from n = k1); nothing flows through the sign product, through a
//     zero input or through the 1e10 stand-in (layers.py:48-58, torch.min / abs / sign backward);
//     Variable + Residual backward (per base column, lane-local): g_c2v[n] = sum_{e != n} gx[e]; g_w_ch[e] += gx[e] llr[e];
//     g_w_res[i] += gx . x_{l-1-i};  gx_{l-1-i} += w_res[i] gx   (only where that output is a queue entry);
//   level 0 (x_0 = llr_e) owns no parameter and is skipped.
// State: three gradient arrays (gx of this level and the pass-through of the next two) in Tensor Memory, g_c2v in shared
// memory, g_w_ch accumulated over the CTA's codewords in shared memory and flushed once with global atomics.
template <class BG>
constexpr int nqb_group_floats() { return BG::kCoreEdges * 32 + 4 * BG::kExtCols * 32; }
template <class BG>
constexpr size_t neural_qc_bwd_smem_bytes() {
    return sizeof(float) * ((size_t)BG::kEdges * 32 + (size_t)kNqGroups * nqb_group_floats<BG>());
}

template <class BG>
__global__ void __launch_bounds__(kNqThreads, 1) neural_qc_bwd_kernel(const NeuralQcBwdParams p) {
    constexpr int Z = BG::kZ, G = 32 / Z;                                  // G codewords per warp, as in the forward kernel
    static_assert(Z == 32 || Z == 16 || Z == 8 || Z == 4, "lifting sizes with a whole number of codewords per warp");
    constexpr int EB = BG::kEdges, EC = BG::kCoreEdges, NX = BG::kExtCols;
    constexpr int ECP = (EC + 3) / 4 * 4;
    constexpr int E = EB * Z;
    extern __shared__ float nq_smem[];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int grp = warp & (kNqGroups - 1), mem = warp / kNqGroups;
    const int r = lane & (Z - 1), sub = lane / Z, lbase = lane & ~(Z - 1);
    auto rot = [&](float v, int s) {
        if constexpr (Z == 32) return __shfl_sync(kFull, v, lane + s);
        else return __shfl_sync(kFull, v, lbase | ((lane + s) & (Z - 1)));
    };
    auto rot_back = [&](float v, int s) {
        if constexpr (Z == 32) return __shfl_sync(kFull, v, lane - s);
        else return __shfl_sync(kFull, v, lbase | ((lane - s) & (Z - 1)));
    };
    float* gw = nq_smem;                                                   // [EB][32]  g_w_ch per (cell, lane), all codewords
    float* gc = nq_smem + EB * 32 + grp * nqb_group_floats<BG>();         // [EC][32]  d/d c2v of the core cells (variable-aligned)
    float* ae = gc + EC * 32;                                              // [3][NX][32] gradient ring of the degree-1 cells
    float* gce = ae + 3 * NX * 32;                                         // [NX][32]  d/d c2v of the degree-1 cells (last level only)

    __shared__ uint32_t tmem_base_s;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < EB * 32; i += kNqThreads) gw[i] = 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = __shfl_sync(kFull, tmem_base_s, 0) + (((uint32_t)grp * 32u) << 16);
    uint32_t tA0 = tbase, tA1 = tbase + ECP, tA2 = tbase + 2 * ECP;       // gx of this level | pass-through to l-1 | to l-2
    int a0 = 0, a1 = NX * 32, a2 = 2 * NX * 32;
    const float wres0 = p.L >= 1 ? __ldg(p.w_res) : 0.0f, wres1 = p.L >= 2 ? __ldg(p.w_res + 1) : 0.0f;
    float acc_wr0 = 0.0f, acc_wr1 = 0.0f;

    for (long long cw0 = (long long)blockIdx.x * kNqGroups * G; cw0 < p.B; cw0 += (long long)gridDim.x * kNqGroups * G) {
        const bool live = cw0 + grp * G + sub < p.B;
        const long long cw = live ? cw0 + grp * G + sub : p.B - 1;
        // the NEXT codeword of this group: pull its saved activations (all levels) into L2 -- they stream from HBM otherwise
        // and every row / column step would wait a full DRAM latency (first build: long_scoreboard 10 stalls per issue)
        {
            const long long nxt = cw0 + ((long long)gridDim.x * kNqGroups + grp) * G;
            if (nxt < p.B) {
                const int t = mem * 32 + lane;
                const long long have = (p.B - nxt < G ? p.B - nxt : G) * (long long)E;     // the G codewords are adjacent
                for (int l = 0; l < p.iters; ++l) {
                    const float* base = p.save_x + ((long long)l * p.B + nxt) * E;
                    for (long long i = 32 * t; i < have; i += 32 * nq::kMembers * 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(base + i));
                }
            }
        }
        // ---- OutputLayer backward at the arg-max edge ----
        const int es = __ldg(p.argmax + cw);
        float gz;
        {
            const float s = p.star ? __ldg(p.star + 2 * cw) : __ldg(p.soft + cw * E + es);
            const float y = p.star ? __ldg(p.star + 2 * cw + 1) : __ldg(p.gt + cw * E + es);
            gz = __ldg(p.g_ml + cw) * (s - y) / fmaxf((1.0f - s) * s, 1e-12f) * (s * (1.0f - s));
            if (!live) gz = 0.0f;                                          // a group past the end of the batch contributes nothing
        }
        const unsigned smeta = nq::chunk_meta[es / Z];
        const int sD = smeta & 0xff, sd = (smeta >> 8) & 0x1f;
        const int soff = Z * ((es / Z) - sD) + (es & (Z - 1)), srow = (soff * (int)(smeta >> 13)) >> 16;      // circulant row of e*
        for (int m = mem; m < EC; m += nq::kMembers) {
            gc[m * 32 + lane] = (m >= sD && m < sD + sd && r == srow) ? gz : 0.0f;
            nq_st1(tA0 + m, 0.0f);
            nq_st1(tA1 + m, 0.0f);
        }
        for (int x = mem; x < NX; x += nq::kMembers) {
            gce[x * 32 + lane] = (EC + x == sD && r == srow) ? gz : 0.0f;
            ae[a0 + x * 32 + lane] = 0.0f;
            ae[a1 + x * 32 + lane] = 0.0f;
        }
        nq_group_sync(grp);

        // Saved activations stream from HBM / L2: every step keeps its loads in flight ahead of their use.  Rows: the 10 + 1
        // values of the member's NEXT row (whatever its shape; unused slots re-read cell 0) are requested before the
        // current row is worked on.
        float pre[11];
        auto row_prefetch = [&](int row, const float* xl) {
            static_for<0, 10>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                pre[k] = __ldg(xl + Z * (int)(nq::row_meta[row][k] & 0xff));
            });
            const int xr = nq::row_ext[row];
            pre[10] = __ldg(xl + Z * (EC + (xr < NX ? xr : 0)));                   // rows without a degree-1 edge re-read slot 0
        };
        auto row_bwd = [&](auto ncc, auto nec, int row, int next_row, const float* xl, bool last) {
            constexpr int NC = decltype(ncc)::value, NE = decltype(nec)::value, d = NC + NE;
            float v[d], g[d], old[NC];
            unsigned sb[d];
            int cell[NC], sft[NC];
            static_for<0, NC>([&](auto kc) { v[decltype(kc)::value] = pre[decltype(kc)::value]; });
            if constexpr (NE) v[NC] = pre[10];
            row_prefetch(next_row, xl);
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const unsigned meta = nq::row_meta[row][k];
                cell[k] = meta & 0xff;
                sft[k] = Z == 32 ? (meta >> 8) : ((meta >> 8) & (Z - 1));
                g[k] = gc[cell[k] * 32 + lane];
                nq_ld1_issue(tA0 + cell[k], old[k]);
            });
            int xs = 0;
            if constexpr (NE) {
                xs = nq::row_ext[row];
                g[NC] = last ? gce[xs * 32 + lane] : 0.0f;      // a degree-1 variable consumes its check message only in `final`
            }
            static_for<0, NC>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                v[k] = rot(v[k], sft[k]);
                g[k] = rot(g[k], sft[k]);
            });
            // forward quantities again: sign bits of (v + 1e-10), first arg-min k1 and runner-up k2 of |v| (zeros -> 1e10)
            unsigned nb = 0;
            float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
            int k1 = -1, k2 = -1;
            bool r1 = false, r2 = false;
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                sb[k] = f2u(__fadd_rn(v[k], 1e-10f));
                nb ^= sb[k];
                const float av = fabsf(v[k]);
                const bool real = av > 0.0f;
                const float a = real ? av : 1e10f;
                if (a < m1) { m2 = m1; k2 = k1; r2 = r1; m1 = a; k1 = k; r1 = real; }
                else if (a < m2) { m2 = a; k2 = k; r2 = real; }
            });
            // fewer than 9 other edges: a minimum above 1e10 loses to the padded slot and carries no gradient
            if (d - 1 < 9) { r1 = r1 && !(m1 > 1e10f); r2 = r2 && !(m2 > 1e10f); }
            float tex = 0.0f, gs1 = 0.0f;                          // sum_{n != k1} g_n sign_n,  g_k1 sign_k1
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const float gs = u2f(f2u(g[k]) ^ ((nb ^ sb[k]) & 0x80000000u));
                tex += (k == k1) ? 0.0f : gs;
                gs1 += (k == k1) ? gs : 0.0f;
            });
            const float val1 = r1 ? tex : 0.0f, val2 = r2 ? gs1 : 0.0f;
            nq_wait_ld();
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                float c = (k == k1) ? val1 : ((k == k2) ? val2 : 0.0f);
                c = u2f(f2u(c) ^ (f2u(v[k]) & 0x80000000u));                       // d|x|/dx = sign(x)
                if constexpr (k < NC) {
                    nq_tie(old[k]);
                    nq_st1(tA0 + cell[k], __fadd_rn(old[k], rot_back(c, sft[k])));
                } else {
                    ae[a0 + xs * 32 + lane] += c;
                }
            });
        };

        for (int l = p.iters - 1; l >= 1; --l) {
            const float* xl = p.save_x + ((long long)l * p.B + cw) * E + r;
            const bool last = l == p.iters - 1;
            row_prefetch((int)nq::sched_rows[mem][0], xl);
            static_for<0, nq::kRowClasses>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const int t1 = nq::sched_row_ptr[mem][c + 1];
#pragma unroll 1
                for (int t = nq::sched_row_ptr[mem][c]; t < t1; ++t)
                    row_bwd(IC<kNqRowNc[c]>{}, IC<kNqRowNe[c]>{}, (int)nq::sched_rows[mem][t],
                            (int)nq::sched_rows[mem][t + 1 < nq::kRowsMax ? t + 1 : t], xl, last);
            });
            nq_group_sync(grp);
            // ---- Variable + Residual backward ----
            const float* x0 = p.save_x + cw * E + r;                                      // llr_e
            const float* xm1 = p.save_x + ((long long)(l - 1) * p.B + cw) * E + r;         // x_{l-1}: queue entry iff l-1 >= 1
            const float* xm2 = p.save_x + ((long long)(l >= 2 ? l - 2 : 0) * p.B + cw) * E + r;
            const float w0 = l >= 2 ? wres0 : 0.0f, w1 = l >= 3 ? wres1 : 0.0f;
            const float u0 = l >= 2 ? 1.0f : 0.0f, u1 = l >= 3 ? 1.0f : 0.0f;            // which w_res entries this level feeds
            // column sums of gx (the member owns at most kColsMax columns; their sums stay in registers)
            float S[nq::kColsMax];
            const int c1 = nq::sched_col_ptr[mem][nq::kColClasses];
            static_for<0, nq::kColsMax>([&](auto sc) {
                constexpr int sl = decltype(sc)::value;
                S[sl] = 0.0f;
                if (sl < c1) {
                    const int j = nq::sched_cols[mem][sl], b0 = nq::col_b0[j], d = nq::col_d[j];
#pragma unroll 1
                    for (int k0 = 0; k0 < d; k0 += 8) {
                        float gx[8];
                        static_for<0, 8>([&](auto ic) { nq_ld1_issue(tA0 + b0 + k0 + decltype(ic)::value, gx[decltype(ic)::value]); });
                        nq_wait_ld();
                        static_for<0, 8>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            nq_tie(gx[i]);
                            S[sl] += (k0 + i < d) ? gx[i] : 0.0f;
                        });
                    }
                }
            });
            // the member's cells, eight at a time: 24 global loads + 16 TMEM loads in flight per chunk
            const int ncell = nq::sched_cell_cnt[mem];
#pragma unroll 1
            for (int c0 = 0; c0 < ncell; c0 += 8) {
                float gx[8], pa[8], ll[8], y1[8], y2[8];
                int cm[8];
                static_for<0, 8>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    const int m = nq::sched_cells[mem][c0 + i];
                    cm[i] = m < 255 ? m : 0;                                        // padding: reads cell 0, stores nothing
                    ll[i] = __ldg(x0 + Z * cm[i]);
                    y1[i] = __ldg(xm1 + Z * cm[i]);
                    y2[i] = __ldg(xm2 + Z * cm[i]);
                    nq_ld1_issue(tA0 + cm[i], gx[i]);
                    nq_ld1_issue(tA1 + cm[i], pa[i]);
                });
                nq_wait_ld();
                static_for<0, 8>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    nq_tie(gx[i]); nq_tie(pa[i]);
                    if (nq::sched_cells[mem][c0 + i] < 255) {
                        const int sl = nq::sched_cell_slot[mem][c0 + i];
                        const float Ss = sl == 0 ? S[0] : (sl == 1 ? S[1] : (sl == 2 ? S[2] : S[3]));
                        gc[cm[i] * 32 + lane] = Ss - gx[i];
                        atomicAdd(&gw[cm[i] * 32 + lane], gx[i] * ll[i]);
                        acc_wr0 = fmaf(u0 * gx[i], y1[i], acc_wr0);
                        acc_wr1 = fmaf(u1 * gx[i], y2[i], acc_wr1);
                        nq_st1(tA1 + cm[i], __fadd_rn(pa[i], w0 * gx[i]));
                        nq_st1(tA2 + cm[i], w1 * gx[i]);
                    }
                });
            }
            const int nx = nq::sched_ext_cnt[mem];
#pragma unroll 1
            for (int t = 0; t < nx; t += 4) {
                float ll[4], y1[4], y2[4];
                int xs[4];
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    xs[i] = nq::sched_ext[mem][t + i < nx ? t + i : t];
                    ll[i] = __ldg(x0 + Z * (EC + xs[i]));
                    y1[i] = __ldg(xm1 + Z * (EC + xs[i]));
                    y2[i] = __ldg(xm2 + Z * (EC + xs[i]));
                });
                static_for<0, 4>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    if (t + i < nx) {
                        const float gx = ae[a0 + xs[i] * 32 + lane];
                        atomicAdd(&gw[(EC + xs[i]) * 32 + lane], gx * ll[i]);
                        acc_wr0 = fmaf(u0 * gx, y1[i], acc_wr0);
                        acc_wr1 = fmaf(u1 * gx, y2[i], acc_wr1);
                        ae[a1 + xs[i] * 32 + lane] += w0 * gx;
                        ae[a2 + xs[i] * 32 + lane] = w1 * gx;
                    }
                });
            }
            nq_group_sync(grp);
            { const uint32_t t = tA0; tA0 = tA1; tA1 = tA2; tA2 = t; }
            { const int t = a0; a0 = a1; a1 = a2; a2 = t; }
        }
        nq_group_sync(grp);
    }
    // ---- flush the parameter gradients ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_wr0 += __shfl_xor_sync(kFull, acc_wr0, o);
        acc_wr1 += __shfl_xor_sync(kFull, acc_wr1, o);
    }
    if (lane == 0) {
        if (p.L >= 1) atomicAdd(p.g_wres, acc_wr0);
        if (p.L >= 2) atomicAdd(p.g_wres + 1, acc_wr1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    for (int m = warp; m < EB; m += kNqThreads / 32) {
        const unsigned cmeta = nq::chunk_meta[m];
        const int D = cmeta & 0xff, d = (cmeta >> 8) & 0x1f;
        atomicAdd(p.g_wch + Z * D + r * d + (m - D), gw[m * 32 + lane]);     // Z < 32: the G codeword slots of a warp add to the same weight
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

}  // namespace ldpc
