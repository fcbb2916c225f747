// neural_qc.cuh -- LDPCNeuralDecoder on the quasi-cyclic structure (5G BG2, Z = 32): no index tables at all.
//
// Same arithmetic, in the same order, as the generic one-kernel decoder (neural.cuh) and therefore as the per-layer
// chain CheckLayer -> VariableLayer -> ResidualLayer -> OutputLayer (models/layers.py:14-66, 78-125, 143-168,
// 180-210 composed as models/decoder.py documents) -- the outputs are bit-identical -- but the neighbour lists the
// reference materialises as (E,9) / (E,22) int64 tables are implied by the base graph:
//   * warp <-> one codeword, lane <-> circulant row; the other edges of a CHECK are the other cells of the base row,
//     reached by one __shfl_sync rotation each (lane + shift), the other edges of a VARIABLE are the other cells of
//     the base column, which live in the SAME lane once messages are kept variable-aligned;
//   * the edge-space state of a codeword -- c2v and a two-deep ring of variable outputs x (159 core cells x 3) --
//     lives in Tensor Memory (tcgen05.ld/st.32x32b: lane <-> TMEM lane, cell <-> column; 480 of the warp's 512
//     columns), the per-edge channel LLRs and the 38 degree-1 cells in shared memory, the per-edge weights w_ch
//     (25 KB, shared by the CTA's warps) in shared memory as [cell][lane];
//   * phase A (per base row): rotate x into check alignment, sign product / two smallest magnitudes with the
//     layer's zero rules, rotate the results back, store c2v;  phase B (per base column, lane-local): the sums
//     over the OTHER edges in the reference's order (running prefix + suffix chain), w_ch * llr + sum + residuals;
//   * the caller's edge numbering is create_LLR_mapping's (utils/ldpc_utils.py:62-95): variable-major,
//     e = 32*D_j + r*d_j + k for column j (D_j base edges before it, degree d_j), circulant row r, k-th check of
//     the variable.  Rows of llr_e / soft are moved with coalesced 128-byte accesses and transposed through a
//     pitch-33 shared-memory tile.
// The Python layer (models/decoder.py) takes this path only after checking that the caller's index tensors ARE
// those tables; anything else runs the table-driven kernels.
#pragma once
#include <math_constants.h>

#include "bg2_tables.h"
#include "params.cuh"

namespace ldpc {

struct NeuralQcParams {
    const float* llr;      // [B, E] edge-space channel LLRs (variable-major edge order)
    const float* w_ch;     // [E]
    const float* w_res;    // [L]
    int L;                 // residual depth, 0..2
    int iters;
    long long B;
    const float* gt;       // [B, E] or null
    float* soft;           // [B, E]
    float* max_loss;       // [B] (with gt)
    // training forward: activations the backward pass reads, or null
    float* save_x;         // [iters, B, E] input of every CheckLayer (x_0 = llr_e, x_1, ...)
};

constexpr int kNqWarps = 4;       // one per TMEM lane quarter: a warp owns all 512 columns of its 32 lanes
constexpr int kNqPitch = 33;      // transposing tile: conflict-free [cell][lane] reads, <= 2-way conflicts on the row side

template <class BG>
constexpr int nq_col_of_vm(int m) {       // base column that owns variable-major base edge m
    int j = 0;
    while (j + 1 < BG::kCols && BG::col_vm0[j + 1] <= m) ++j;
    return j;
}
template <class BG>
constexpr size_t neural_qc_smem_bytes() {
    return sizeof(float) * ((size_t)BG::kEdges * 32 + (size_t)kNqWarps * (BG::kEdges * kNqPitch + 3 * BG::kExtCols * 32));
}

__device__ __forceinline__ void nq_ld1_issue(uint32_t taddr, float& a) {
    uint32_t x;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr));
    a = u2f(x);
}
__device__ __forceinline__ void nq_ld4_issue(uint32_t taddr, float& a, float& b, float& c, float& d) {
    uint32_t x, y, z, w;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr));
    a = u2f(x); b = u2f(y); c = u2f(z); d = u2f(w);
}
__device__ __forceinline__ void nq_tie(float& a) { asm volatile("" : "+f"(a)); }
__device__ __forceinline__ void nq_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(f2u(v)) : "memory");
}
__device__ __forceinline__ void nq_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(f2u(a)), "r"(f2u(b)), "r"(f2u(c)), "r"(f2u(d)) : "memory");
}
__device__ __forceinline__ void nq_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void nq_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// defined in neural_qc.cu
int launch_neural_qc(const ldpc_code_t* c, const NeuralQcParams& p, cudaStream_t st);

template <class BG>
__global__ void __launch_bounds__(kNqWarps * 32, 1) neural_qc_kernel(const NeuralQcParams p) {
    static_assert(BG::kZ == 32, "one codeword per warp");
    constexpr int Z = 32, EB = BG::kEdges, EC = BG::kCoreEdges, NX = BG::kExtCols, NC = BG::kCoreCols;
    constexpr int EQ = (EC + 3) / 4, ECP = EQ * 4;         // quads / padded core cells per state array
    constexpr int E = EB * Z;
    static_assert(3 * ECP <= 512, "c2v + two ring slots must fit the warp's 512 TMEM columns");
    static_assert(BG::core_cm0[NC] == EC, "core columns come first in the variable-major numbering");
    extern __shared__ float nq_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* wsm = nq_smem;                                                    // [EB][32]  w_ch per (cell, lane)
    float* lls = nq_smem + EB * 32 + warp * (EB * kNqPitch + 3 * NX * 32);   // [EB][33]  llr_e, later the soft outputs
    float* xe0 = lls + EB * kNqPitch;                                        // [2][NX][32] ring of the degree-1 cells
    float* ces = xe0 + 2 * NX * 32;                                          // [NX][32]  their last check message

    __shared__ uint32_t tmem_base_s;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // w_ch[e] -> wsm[cell][lane], e = 32*D_j + lane*d_j + k
    static_for<0, EB>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        if (m % kNqWarps == warp) {
            constexpr int j = nq_col_of_vm<BG>(m), d = BG::col_deg[j], D = BG::col_vm0[j];
            wsm[m * 32 + lane] = __ldg(p.w_ch + 32 * D + lane * d + (m - D));
        }
    });
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s + (((uint32_t)warp * 32u) << 16);
    const uint32_t tC = tbase;                        // c2v
    uint32_t tXc = tbase + ECP, tXo = tbase + 2 * ECP; // ring: current x, older x (roles swap every iteration)
    int xc_off = 0, xo_off = NX * 32;                  // the same for the degree-1 ring in shared memory
    const float wres0 = p.L >= 1 ? __ldg(p.w_res) : 0.0f, wres1 = p.L >= 2 ? __ldg(p.w_res + 1) : 0.0f;

    int lp[Z];
    static_for<0, Z>([&](auto tc) {
        constexpr int t = decltype(tc)::value;
        asm("add.s32 %0, %1, %2;" : "=r"(lp[t]) : "r"(lane), "n"(t));
    });

    for (long long cw = (long long)blockIdx.x * kNqWarps + warp; cw < p.B; cw += (long long)gridDim.x * kNqWarps) {
        // ---- load llr_e[cw]: coalesced rows -> [cell][lane] tile ----
        {
            const float* src = p.llr + cw * E + lane;
            static_for<0, EB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                constexpr int j = nq_col_of_vm<BG>(m), d = BG::col_deg[j], D = BG::col_vm0[j];
                const float v = __ldg(src + 32 * m);
                const int off = 32 * (m - D) + lane, r = off / d, k = off - r * d;
                lls[(D + k) * kNqPitch + r] = v;
            });
        }
        __syncwarp();
        // x_0 = llr_e in the current ring slot, zeros in the older one (its residual weight is zero until it is written)
        static_for<0, EQ>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            float v[4];
            static_for<0, 4>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                v[i] = (4 * q + i < EC) ? lls[(4 * q + i) * kNqPitch + lane] : 0.0f;
            });
            nq_st4(tXc + 4 * q, v[0], v[1], v[2], v[3]);
            nq_st4(tXo + 4 * q, 0.f, 0.f, 0.f, 0.f);
        });
        static_for<0, NX>([&](auto xc) {
            constexpr int x = decltype(xc)::value;
            xe0[xc_off + x * 32 + lane] = lls[(EC + x) * kNqPitch + lane];
            xe0[xo_off + x * 32 + lane] = 0.0f;
        });
        nq_wait_st();

        // ---- phase A: CheckLayer on the current x, one base row at a time ----
        auto phase_a = [&](auto lastc) {
            constexpr bool kLast = decltype(lastc)::value != 0;
            constexpr int kGrp = 2, kNumGrp = (BG::kRows + kGrp - 1) / kGrp;
            float xv[EB];
            auto for_group_core = [&](auto gc, auto&& fn) {
                constexpr int g = decltype(gc)::value;
                constexpr int ea = BG::row_ptr[g * kGrp], eb = BG::row_ptr[(g + 1) * kGrp < BG::kRows ? (g + 1) * kGrp : BG::kRows];
                static_for<ea, eb>([&](auto ec) {
                    if constexpr (BG::kind[decltype(ec)::value] == 0) fn(ec);
                });
            };
            auto issue = [&](auto gc) {
                for_group_core(gc, [&](auto ec) {
                    constexpr int e = decltype(ec)::value, cme = BG::cm[e];
                    nq_ld1_issue(tXc + cme, xv[e]);
                });
            };
            auto settle = [&](auto gc) {
                nq_wait_ld();
                for_group_core(gc, [&](auto ec) { nq_tie(xv[decltype(ec)::value]); });
            };
            issue(IC<0>{});
            static_for<0, BG::kRows>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int e0 = BG::row_ptr[i], d = BG::row_ptr[i + 1] - e0;
                if constexpr (i % 2 == 0 && i > 0) asm volatile("" ::: "memory");
                if constexpr (i % kGrp == 0) {
                    settle(IC<i / kGrp>{});
                    if constexpr (i / kGrp + 1 < kNumGrp) issue(IC<i / kGrp + 1>{});
                }
                float v[d], a[d];
                unsigned sb[d];
                int zi[d];
                unsigned nb = 0;
                int zc = 0;
                float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0) {
                        constexpr int s = BG::shift[e];
                        v[k] = (s == 0) ? xv[e] : __shfl_sync(kFull, xv[e], lp[s], Z);
                    } else {
                        static_assert(BG::kind[e] == 0 || BG::shift[e] == 0, "degree-1 columns are expected unshifted");
                        constexpr int xs = BG::slot[e];
                        v[k] = xe0[xc_off + xs * 32 + lane];
                    }
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
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, e = e0 + k;
                    if constexpr (BG::kind[e] == 0 || kLast) {
                        const float m = (a[k] == m1) ? m2c : m1c;                       // minimum over the OTHER edges
                        const unsigned sgn = (nb ^ sb[k]) & 0x80000000u;               // their sign product
                        const float sp = u2f(sgn | ((zc - zi[k]) > 0 ? 0u : 0x3f800000u));
                        const float o = __fmul_rn(sp, m);
                        if constexpr (BG::kind[e] == 0) {
                            constexpr int s = BG::shift[e], cme = BG::cm[e];
                            const float t = (s == 0) ? o : __shfl_sync(kFull, o, lp[Z - s], Z);
                            nq_st1(tC + cme, t);
                        } else {
                            constexpr int xs = BG::slot[e];
                            ces[xs * 32 + lane] = o;
                        }
                    }
                });
            });
            nq_wait_st();
        };

        // sums over the OTHER edges of a variable, in the table's (ascending check) order, starting from 0
        // (neural_gather_sum): running prefix + suffix chain = the same fp32 additions
        auto column_sums = [&](auto cc, const float* cv, float* acc) {
            constexpr int c = decltype(cc)::value, b0 = BG::core_cm0[c], d = BG::core_cm0[c + 1] - b0;
            static_assert(d <= 32, "column degree");
            float pre = 0.0f;
            static_for<0, d>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                float s = pre;
                static_for<k + 1, d>([&](auto k2) { s = __fadd_rn(s, cv[b0 + decltype(k2)::value]); });
                acc[k] = s;
                pre = __fadd_rn(pre, cv[b0 + k]);
            });
        };
        // quads of the state arrays are fetched one base column ahead; a quad belongs to the column of its first cell
        auto for_col_quads = [&](auto cc, auto&& fn) {
            constexpr int c = decltype(cc)::value;
            constexpr int qa = (BG::core_cm0[c] + 3) / 4, qb = (BG::core_cm0[c + 1] + 3) / 4;
            static_for<qa, qb>([&](auto qc) { fn(qc); });
        };

        for (int l = 0; l < p.iters; ++l) {
            if (p.save_x) {
                // training forward: the input of this CheckLayer, in the caller's edge order (coalesced via the tile would
                // need a second staging area; the values are re-read lane-locally and stored with the d_j-strided pattern)
                float* dst = p.save_x + ((long long)l * p.B + cw) * E;
                float xq[ECP];
                static_for<0, EQ>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    nq_ld4_issue(tXc + 4 * q, xq[4 * q], xq[4 * q + 1], xq[4 * q + 2], xq[4 * q + 3]);
                });
                nq_wait_ld();
                static_for<0, EC>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    constexpr int j = nq_col_of_vm<BG>(m), d = BG::col_deg[j], D = BG::col_vm0[j];
                    nq_tie(xq[m]);
                    dst[32 * D + lane * d + (m - D)] = xq[m];
                });
                static_for<0, NX>([&](auto xc) {
                    constexpr int x = decltype(xc)::value;
                    dst[32 * (EC + x) + lane] = xe0[xc_off + x * 32 + lane];
                });
            }
            if (l == p.iters - 1) {
                phase_a(IC<1>{});
                break;
            }
            phase_a(IC<0>{});
            // ---- phase B: VariableLayer(0, c2v) + ResidualLayer, one base column at a time, lane-local ----
            // queue of earlier outputs (models/decoder.py): x_0 is not an entry, so the first update has no residual term
            const float wr0 = l >= 1 ? wres0 : 0.0f, wr1 = l >= 2 ? wres1 : 0.0f;
            {
                float cv[ECP], xc[ECP], xo[ECP], nx[ECP];
                auto issue = [&](auto cc) {
                    for_col_quads(cc, [&](auto qc) {
                        constexpr int b = decltype(qc)::value * 4;
                        nq_ld4_issue(tC + b, cv[b], cv[b + 1], cv[b + 2], cv[b + 3]);
                        nq_ld4_issue(tXc + b, xc[b], xc[b + 1], xc[b + 2], xc[b + 3]);
                        nq_ld4_issue(tXo + b, xo[b], xo[b + 1], xo[b + 2], xo[b + 3]);
                    });
                };
                auto settle = [&](auto cc) {
                    nq_wait_ld();
                    for_col_quads(cc, [&](auto qc) {
                        constexpr int b = decltype(qc)::value * 4;
                        static_for<0, 4>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            nq_tie(cv[b + i]); nq_tie(xc[b + i]); nq_tie(xo[b + i]);
                        });
                    });
                };
                issue(IC<0>{});
                static_for<0, NC>([&](auto cc) {
                    constexpr int c = decltype(cc)::value, b0 = BG::core_cm0[c], d = BG::core_cm0[c + 1] - b0;
                    asm volatile("" ::: "memory");
                    settle(cc);
                    if constexpr (c + 1 < NC) issue(IC<c + 1>{});
                    float acc[32];
                    column_sums(cc, cv, acc);
                    static_for<0, d>([&](auto kc) {
                        constexpr int k = decltype(kc)::value, m = b0 + k;
                        const float ll = lls[m * kNqPitch + lane], w = wsm[m * 32 + lane];
                        float r = __fadd_rn(__fmul_rn(ll, w), acc[k]);
                        r = __fadd_rn(r, __fmul_rn(wr0, xc[m]));
                        r = __fadd_rn(r, __fmul_rn(wr1, xo[m]));
                        nx[m] = r;
                        if constexpr (m % 4 == 3 || m == EC - 1) {
                            constexpr int b = (m / 4) * 4;
                            nq_st4(tXo + b, nx[b], b + 1 < EC ? nx[b + 1] : 0.f, b + 2 < EC ? nx[b + 2] : 0.f, b + 3 < EC ? nx[b + 3] : 0.f);
                        }
                    });
                });
                static_for<0, NX>([&](auto xcn) {
                    constexpr int x = decltype(xcn)::value, m = EC + x;
                    const float ll = lls[m * kNqPitch + lane], w = wsm[m * 32 + lane];
                    float r = __fadd_rn(__fmul_rn(ll, w), 0.0f);                    // a degree-1 variable has no other edge
                    r = __fadd_rn(r, __fmul_rn(wr0, xe0[xc_off + x * 32 + lane]));
                    r = __fadd_rn(r, __fmul_rn(wr1, xe0[xo_off + x * 32 + lane]));
                    xe0[xo_off + x * 32 + lane] = r;
                });
                nq_wait_st();
            }
            { const uint32_t t = tXc; tXc = tXo; tXo = t; }
            { const int t = xc_off; xc_off = xo_off; xo_off = t; }
        }

        // ---- final = VariableLayer(c2v, c2v); OutputLayer: soft = sigmoid(final + llr) ----
        {
            float cv[ECP];
            static_for<0, EQ>([&](auto qc) {
                constexpr int b = decltype(qc)::value * 4;
                nq_ld4_issue(tC + b, cv[b], cv[b + 1], cv[b + 2], cv[b + 3]);
            });
            nq_wait_ld();
            static_for<0, ECP>([&](auto mc) { nq_tie(cv[decltype(mc)::value]); });
            static_for<0, NC>([&](auto cc) {
                constexpr int c = decltype(cc)::value, b0 = BG::core_cm0[c], d = BG::core_cm0[c + 1] - b0;
                float acc[32];
                column_sums(cc, cv, acc);
                static_for<0, d>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, m = b0 + k;
                    const float ll = lls[m * kNqPitch + lane];
                    const float z = __fadd_rn(__fadd_rn(cv[m], acc[k]), ll);
                    lls[m * kNqPitch + lane] = 1.0f / (1.0f + expf(-z));
                });
            });
            static_for<0, NX>([&](auto xcn) {
                constexpr int x = decltype(xcn)::value, m = EC + x;
                const float ll = lls[m * kNqPitch + lane];
                const float z = __fadd_rn(__fadd_rn(ces[x * 32 + lane], 0.0f), ll);
                lls[m * kNqPitch + lane] = 1.0f / (1.0f + expf(-z));
            });
        }
        __syncwarp();
        {
            float* dst = p.soft + cw * E + lane;
            const float* gts = p.gt ? p.gt + cw * E + lane : nullptr;
            float best = -CUDART_INF_F;
            static_for<0, EB>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                constexpr int j = nq_col_of_vm<BG>(m), d = BG::col_deg[j], D = BG::col_vm0[j];
                const int off = 32 * (m - D) + lane, r = off / d, k = off - r * d;
                const float s = lls[(D + k) * kNqPitch + r];
                dst[32 * m] = s;
                if (gts) {
                    const float y = __ldg(gts + 32 * m);
                    const float l1 = fmaxf(logf(s), -100.0f), l0 = fmaxf(logf(1.0f - s), -100.0f);
                    const float loss = -(y * l1 + (1.0f - y) * l0);
                    best = loss > best ? loss : best;
                }
            });
            if (gts) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(kFull, best, o));
                if (lane == 0) p.max_loss[cw] = best;
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

}  // namespace ldpc
