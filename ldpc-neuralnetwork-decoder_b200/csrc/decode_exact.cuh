// decode_exact.cuh -- table-driven flooding decoder, reference operation order.
//
// Replaces MinSumScaledDecoder.decode (models/traditional_decoders.py:177-260) and
// BeliefPropagationDecoder.decode (:42-109) for ANY quasi-cyclic code with Z <= 32
// (LDPC_PATH_EXACT).  It reproduces the reference's fp32 arithmetic operation for operation:
//   check update   min-sum: sign product and min over the OTHER edges of the check, then
//                  alpha*min (fp32), then sign apply                       (:207-232)
//                  BP: prod of tanh(v/2) over the other edges in ascending variable order,
//                  2*atanh(prod), unclipped (inf/NaN propagate)             (:72-81)
//   variable update llr + c2v of the other checks, added in ascending check order (:235-244)
//   posterior      llr + all c2v in ascending check order; bit = belief < 0   (:247-252)
// so min-sum results are bit-identical to the reference and BP results differ only through
// the 1-ulp differences of tanh/atanh (math_ref.cuh).
//
// Mapping: one warp owns G = 32/Z codewords; lane = cw_in_warp*Z + r.  All edge messages of
// the warp's codewords stay in shared memory for every iteration:
//   msg[e][lane]  message on base edge e, CHECK-aligned (lane r <-> check i*Z+r); holds v2c
//                 at the start of the check phase and c2v after it
//   T[j][lane]    posterior of variable j*Z+r', VARIABLE-aligned
//   L[j][lane]    channel LLR, VARIABLE-aligned
// A circulant shift is an address rotation inside the codeword's Z-lane window, so all
// shared-memory accesses of a warp hit 32 distinct banks.  Exclusion sums/products use a
// running prefix (shared) plus a per-edge suffix chain, which is the reference's order.
#pragma once
#include <math_constants.h>

#include "params.cuh"
#include "math_ref.cuh"
#include "tables.cuh"

namespace ldpc {

template <bool kConst>
__device__ __forceinline__ bool syndrome_bad(const Tab<kConst>& tab, const float* T, int rows, int Z, int base, int r) {
    const int off_rowptr = tab[7], off_redge = tab[9];
    unsigned bad = 0;
    for (int i = 0; i < rows; ++i) {
        const int e0 = tab[off_rowptr + i], e1 = tab[off_rowptr + i + 1];
        unsigned par = 0;
        for (int e = e0; e < e1; ++e) {
            const uint32_t w = tab[off_redge + e];
            int rr = r + (int)((w >> 16) & 0xff);
            rr -= rr >= Z ? Z : 0;
            par ^= (T[(w & 0xffff) * 32 + base + rr] < 0.0f) ? 1u : 0u;
        }
        bad |= par;
    }
    return bad != 0;
}

template <int kAlgo, int kMaxDc, int kMaxDv, bool kConst, bool kScratch>
__global__ void __launch_bounds__(256) decode_exact_kernel(const DecodeParams p, float* __restrict__ scratch) {
    // scratch != nullptr: the per-warp state (messages, posteriors, LLRs) does not fit shared memory (large lifting factors
    // held as many small circulants, large Z = 1 codes) and lives in a global workspace instead -- same code, L2 speed
    extern __shared__ float smem[];
    const Tab<kConst> tab{p.gtab, p.slot};
    const int rows = tab[0], cols = tab[1], Z = tab[2], E = tab[3], G = tab[4];
    const int off_rowptr = tab[7], off_colptr = tab[8], off_redge = tab[9], off_cedge = tab[10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    // (a template parameter, not a run-time choice: with one pointer for both address spaces every access became a generic
    // LD / ST instead of LDS / STS)
    float* msg;
    if constexpr (kScratch) msg = scratch + ((size_t)blockIdx.x * W + warp) * p.floats_per_warp;
    else msg = smem + (size_t)warp * p.floats_per_warp;
    float* T = msg + E * 32;
    float* L = T + cols * 32;
    const bool active = lane < G * Z;
    const int cwi = active ? lane / Z : 0;
    const int r = active ? lane - cwi * Z : 0;
    const int base = cwi * Z;
    const int N = cols * Z;
    const unsigned gmask = (Z == 32) ? 0xffffffffu : (((1u << Z) - 1u) << base);
    const bool track = p.stop_mode != LDPC_STOP_FIXED || p.valid_mask != nullptr;
    unsigned long long acc_bits = 0, acc_fe = 0, acc_frames = 0, acc_und = 0;

    for (long long grp = (long long)blockIdx.x * W + warp; grp < p.ngroups; grp += (long long)gridDim.x * W) {
        const long long cw = grp * G + cwi;
        const bool live = active && cw < p.B;
        if (p.gen.enabled) {
            for (int j = 0; j < cols; ++j)
                L[j * 32 + lane] = live ? gen_llr(p.gen, p.gen.first_frame + (unsigned long long)cw, j * Z + r) : 0.0f;
        } else {
            const float* llr = p.llr + cw * N + r;
            for (int j = 0; j < cols; ++j) L[j * 32 + lane] = live ? __ldg(llr + j * Z) : 0.0f;
        }
        __syncwarp();
        // v2c := llr on every edge (traditional_decoders.py:64-67,199-202)
        for (int e = 0; e < E; ++e) {
            const uint32_t w = tab[off_redge + e];
            int rr = r + (int)((w >> 16) & 0xff);
            rr -= rr >= Z ? Z : 0;
            msg[e * 32 + lane] = L[(w & 0xffff) * 32 + base + rr];
        }
        for (int j = 0; j < cols; ++j) T[j * 32 + lane] = L[j * 32 + lane];
        __syncwarp();

        bool done = !live;
        int my_iters = p.iters;
        unsigned long long vm = 0;
        for (int it = 0; it < p.iters; ++it) {
            // Both updates are written once for a compile-time bound on the node degree and instantiated for a ladder of
            // bounds; a row / column runs the smallest instance that holds it (warp-uniform choice).  The operations and their
            // order are those of the full-width loops -- the padded slots only ever contributed predicated-off instructions,
            // which at kMaxDv = 24 were 276 additions per base column whatever its degree (38 of BG2's 52 columns have degree 1).
            // ---------------- check-node update ----------------
            auto check_row = [&](auto dcc, int e0, int d) {
                constexpr int DC = decltype(dcc)::value;
                float v[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) v[k] = (k < d) ? msg[(e0 + k) * 32 + lane] : 0.0f;
                if constexpr (kAlgo == LDPC_ALGO_MINSUM) {
                    float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
                    uint32_t sg = 0;
                    int nn = 0;                    // NaN inputs (only reachable from non-finite channel LLRs)
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (k < d) {
                            // reference :216-222: `mag < min_mag` is false for a NaN magnitude, i.e. a NaN never becomes a
                            // minimum: it takes part as +inf here
                            const float a = fabsf(v[k]);
                            const float an = (a != a) ? CUDART_INF_F : a;
                            m2 = fminf(m2, fmaxf(m1, an));
                            m1 = fminf(m1, an);
                            sg ^= f2u(v[k]);
                            nn += (a != a) ? 1 : 0;
                        }
                    const float s1 = __fmul_rn(p.alpha, m1), s2 = __fmul_rn(p.alpha, m2);
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (k < d) {
                            // torch.sign(NaN) = 0 (:213), so a NaN among the OTHER inputs zeroes the sign product: the
                            // message is 0 * scaled_min = 0, or NaN when that minimum is infinite
                            const bool other_nan = nn - ((v[k] != v[k]) ? 1 : 0) > 0;
                            const float mag = (fabsf(v[k]) == m1) ? s2 : s1;
                            const float out = u2f(f2u(mag) ^ ((sg ^ f2u(v[k])) & 0x80000000u));
                            v[k] = other_nan ? __fmul_rn(0.0f, mag) : out;
                        }
                } else {
                    float t[DC];
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        t[k] = 1.0f;
                        if (k < d) t[k] = tanh_half_ref(v[k]);
                    }
                    float pre = 1.0f;
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (k < d) {
                            float pr = pre;
#pragma unroll
                            for (int k2 = k + 1; k2 < DC; ++k2)
                                if (k2 < d) pr = __fmul_rn(pr, t[k2]);
                            v[k] = two_atanh_ref(pr);
                            pre = __fmul_rn(pre, t[k]);
                        }
                }
                if (!done) {
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (k < d) msg[(e0 + k) * 32 + lane] = v[k];
                }
            };
            for (int i = 0; i < rows; ++i) {
                const int e0 = tab[off_rowptr + i];
                const int d = (int)tab[off_rowptr + i + 1] - e0;
                if (d <= 3) check_row(IC<3>{}, e0, d);
                else if (d <= 4) check_row(IC<4>{}, e0, d);
                else if (d <= 5) check_row(IC<5>{}, e0, d);
                else if (d <= 6) check_row(IC<(kMaxDc < 6 ? kMaxDc : 6)>{}, e0, d);
                else if (d <= 8) check_row(IC<(kMaxDc < 8 ? kMaxDc : 8)>{}, e0, d);
                else if (d <= 10) check_row(IC<(kMaxDc < 10 ? kMaxDc : 10)>{}, e0, d);
                else if (d <= 16) check_row(IC<(kMaxDc < 16 ? kMaxDc : 16)>{}, e0, d);
                else check_row(IC<kMaxDc>{}, e0, d);
            }
            __syncwarp();
            // ---------------- variable-node update + posterior ----------------
            auto var_col = [&](auto dvc, int j, int k0, int d) {
                constexpr int DV = decltype(dvc)::value;
                float c[DV];
                int addr[DV];
#pragma unroll
                for (int k = 0; k < DV; ++k) {
                    if (k < d) {
                        const uint32_t w = tab[off_cedge + k0 + k];
                        int rr = r - (int)((w >> 16) & 0xff);
                        rr += rr < 0 ? Z : 0;
                        addr[k] = (w & 0xffff) * 32 + base + rr;
                        c[k] = msg[addr[k]];
                    } else {
                        addr[k] = 0;
                        c[k] = 0.0f;
                    }
                }
                float pre = L[j * 32 + lane];
#pragma unroll
                for (int k = 0; k < DV; ++k)
                    if (k < d) {
                        float s = pre;
#pragma unroll
                        for (int k2 = k + 1; k2 < DV; ++k2)
                            if (k2 < d) s = __fadd_rn(s, c[k2]);
                        pre = __fadd_rn(pre, c[k]);
                        c[k] = s;
                    }
                if (!done) {
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (k < d) msg[addr[k]] = c[k];
                    T[j * 32 + lane] = pre;
                }
            };
            for (int j = 0; j < cols; ++j) {
                const int k0 = tab[off_colptr + j];
                const int d = (int)tab[off_colptr + j + 1] - k0;
                if (d <= 1) var_col(IC<1>{}, j, k0, d);
                else if (d <= 2) var_col(IC<2>{}, j, k0, d);
                else if (d <= 3) var_col(IC<3>{}, j, k0, d);
                else if (d <= 4) var_col(IC<4>{}, j, k0, d);
                else if (d <= 6) var_col(IC<6>{}, j, k0, d);
                else if (d <= 8) var_col(IC<8>{}, j, k0, d);
                else if (d <= 12) var_col(IC<12>{}, j, k0, d);
                else if (d <= 16) var_col(IC<16>{}, j, k0, d);
                else if (d <= 24) var_col(IC<(kMaxDv < 24 ? kMaxDv : 24)>{}, j, k0, d);
                else var_col(IC<kMaxDv>{}, j, k0, d);
            }
            __syncwarp();
            if (track) {
                const bool bad = syndrome_bad(tab, T, rows, Z, base, r);
                const unsigned m = __ballot_sync(0xffffffffu, bad && active);
                const bool ok = (m & gmask) == 0;
                if (ok) vm |= 1ull << (it & 63);
                if (p.valid_mask && live && r == 0 && ((it & 63) == 63 || it == p.iters - 1)) {
                    if ((it >> 6) < p.mask_words) p.valid_mask[cw * p.mask_words + (it >> 6)] = vm;
                }
                if ((it & 63) == 63) vm = 0;
                if (p.stop_mode == LDPC_STOP_PER_CODEWORD) {
                    if (ok && !done) {
                        done = true;
                        my_iters = it + 1;
                        // the frozen codeword stays valid: mark the remaining iterations of this word
                        if (p.valid_mask && live && r == 0) {
                            for (int t2 = it; t2 < p.iters; ++t2)
                                if ((t2 >> 6) < p.mask_words)
                                    atomicOr(&p.valid_mask[cw * p.mask_words + (t2 >> 6)], 1ull << (t2 & 63));
                        }
                    }
                    if (__all_sync(0xffffffffu, done)) break;
                }
            }
        }

        // ---------------- outputs ----------------
        if (p.soft_out && live) {
            float* o = p.soft_out + cw * N + r;
            for (int j = 0; j < cols; ++j) o[j * Z] = T[j * 32 + lane];
        }
        if (p.iters_out && live && r == 0) p.iters_out[cw] = my_iters;
        if (p.syndrome_ok || p.counters) {
            const bool bad = syndrome_bad(tab, T, rows, Z, base, r);
            const unsigned m = __ballot_sync(0xffffffffu, bad && active);
            const bool ok = (m & gmask) == 0;
            if (p.syndrome_ok && live && r == 0) p.syndrome_ok[cw] = ok ? 1 : 0;
            if (p.counters) {
                // all-zero codeword was sent: every negative posterior is a bit error
                unsigned e = 0;
                if (live)
                    for (int j = 0; j < cols; ++j) e += T[j * 32 + lane] < 0.0f ? 1u : 0u;
                unsigned tot = 0;   // errors of this lane's codeword
                for (int l = 0; l < 32; ++l) {
                    const unsigned el = __shfl_sync(0xffffffffu, e, l);
                    if (l >= base && l < base + Z) tot += el;
                }
                if (live && r == 0) {
                    acc_bits += tot;
                    acc_fe += tot != 0;
                    acc_frames += 1;
                    acc_und += (tot != 0 && ok);
                }
            }
        }
        if (p.hard_out) {
            if (p.hard_dtype == LDPC_HARD_F32) {
                float* o = (float*)p.hard_out + cw * N + r;
                if (live)
                    for (int j = 0; j < cols; ++j) o[j * Z] = T[j * 32 + lane] < 0.0f ? 1.0f : 0.0f;
            } else if (p.hard_dtype == LDPC_HARD_U8) {
                uint8_t* o = (uint8_t*)p.hard_out + cw * N + r;
                if (live)
                    for (int j = 0; j < cols; ++j) o[j * Z] = T[j * 32 + lane] < 0.0f ? 1 : 0;
            } else {
                // packed words: stage in the (now free) message area
                const int NW = (N + 31) >> 5;
                unsigned* buf = reinterpret_cast<unsigned*>(msg);
                __syncwarp();
                for (int x = lane; x < G * NW; x += 32) buf[x] = 0;
                __syncwarp();
                if (active)
                    for (int j = 0; j < cols; ++j)
                        if (T[j * 32 + lane] < 0.0f) {
                            const int n = j * Z + r;
                            atomicOr(&buf[cwi * NW + (n >> 5)], 1u << (n & 31));
                        }
                __syncwarp();
                for (int x = lane; x < G * NW; x += 32) {
                    const long long cw2 = grp * G + x / NW;
                    if (cw2 < p.B) ((unsigned*)p.hard_out)[cw2 * NW + x % NW] = buf[x];
                }
            }
        }
        __syncwarp();
    }
    if (p.counters) flush_counters(p.counters, acc_bits, acc_fe, acc_frames, acc_und);
}

// Syndrome of given hard decisions: one warp per G codewords, bits staged as +-1 in shared
// memory (same layout / rotation-by-address as the decoder), parity per check row.
template <bool kConst>
__global__ void __launch_bounds__(256) syndrome_kernel(const uint32_t* gtab, int slot, const void* __restrict__ hard,
                                                        int hard_dtype, long long B, long long ngroups,
                                                        uint8_t* __restrict__ ok_out, float* __restrict__ scratch) {
    extern __shared__ float smem[];
    const Tab<kConst> tab{gtab, slot};
    const int rows = tab[0], cols = tab[1], Z = tab[2], G = tab[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    float* T = scratch ? scratch + ((size_t)blockIdx.x * W + warp) * cols * 32 : smem + (size_t)warp * cols * 32;
    const bool active = lane < G * Z;
    const int cwi = active ? lane / Z : 0, r = active ? lane - cwi * Z : 0, base = cwi * Z, N = cols * Z;
    const unsigned gmask = (Z == 32) ? 0xffffffffu : (((1u << Z) - 1u) << base);
    for (long long grp = (long long)blockIdx.x * W + warp; grp < ngroups; grp += (long long)gridDim.x * W) {
        const long long cw = grp * G + cwi;
        const bool live = active && cw < B;
        for (int j = 0; j < cols; ++j) {
            const long long n = cw * N + j * Z + r;
            bool bit = false;
            if (live) {
                if (hard_dtype == LDPC_HARD_F32) bit = ((const float*)hard)[n] != 0.0f;
                else if (hard_dtype == LDPC_HARD_U8) bit = ((const uint8_t*)hard)[n] != 0;
                else {
                    const int q = j * Z + r;
                    bit = (((const unsigned*)hard)[cw * ((N + 31) >> 5) + (q >> 5)] >> (q & 31)) & 1u;
                }
            }
            T[j * 32 + lane] = bit ? -1.0f : 1.0f;
        }
        __syncwarp();
        const bool bad = syndrome_bad(tab, T, rows, Z, base, r);
        const unsigned m = __ballot_sync(0xffffffffu, bad && active);
        if (live && r == 0) ok_out[cw] = (m & gmask) == 0 ? 1 : 0;
        __syncwarp();
    }
}

// Global workspace for codes whose per-warp state exceeds shared memory: at most ~1 GB, stream-ordered allocation
#ifndef LDPC_EXACT_WS_WARPS
#define LDPC_EXACT_WS_WARPS 16         // warps per SM when the state lives in the global workspace (2 blocks of 8: the register file's limit)
#endif
#ifndef LDPC_EXACT_WS_BELOW
#define LDPC_EXACT_WS_BELOW 4          // min-sum: use the workspace when fewer warps than this fit an SM's shared memory (sum-product: 8)
#endif
struct ExactScratch {
    float* ptr = nullptr;
    int W = 0;
    long long blocks = 0;
    cudaStream_t st = nullptr;
    int init(size_t per_warp, long long ngroups, cudaStream_t stream, const char* who) {
        st = stream;
        long long warps = (long long)((size_t)1 << 30) / (long long)per_warp;
        if (warps < 1) return fail(LDPC_ERR_UNSUPPORTED, "%s: %zu bytes of state per warp", who, per_warp);
        if (warps > (long long)kNumSMs * LDPC_EXACT_WS_WARPS) warps = (long long)kNumSMs * LDPC_EXACT_WS_WARPS;
        if (warps > ngroups) warps = ngroups;
        W = warps < 8 ? (int)warps : 8;
        blocks = warps / W;
        {   // keep the workspace in the device's pool between calls: with the default release threshold (0) every
            // synchronisation hands the memory back to the driver and the next call pays for a fresh allocation
            int dev = 0;
            cudaMemPool_t pool;
            LDPC_CUDA(cudaGetDevice(&dev));
            LDPC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t cur = 0, keep = (uint64_t)2 << 30;
            LDPC_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur));
            if (cur < keep) LDPC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        LDPC_CUDA(cudaMallocAsync((void**)&ptr, per_warp * (size_t)W * (size_t)blocks, st));
        return LDPC_OK;
    }
    ~ExactScratch() { if (ptr) cudaFreeAsync(ptr, st); }
};

inline int launch_syndrome(const ldpc_code* c, const void* hard, int hard_dtype, long long B, uint8_t* ok, cudaStream_t st) {
    const size_t per_warp = (size_t)c->cols * 32 * sizeof(float);
    int W = (int)(kMaxSmemPerBlock / per_warp);
    const long long ngroups = (B + c->G - 1) / c->G;
    ExactScratch sc;
    long long blocks;
    size_t smem = 0;
    if (W < 1) {
        if (int rc = sc.init(per_warp, ngroups, st, "syndrome_check")) return rc;
        W = sc.W; blocks = sc.blocks;
    } else {
        if (W > 8) W = 8;
        blocks = (ngroups + W - 1) / W;
        if (blocks > (long long)kNumSMs * 4) blocks = (long long)kNumSMs * 4;
        smem = per_warp * W;
    }
    if (c->slot >= 0) {
        LDPC_CUDA(cudaFuncSetAttribute(syndrome_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        syndrome_kernel<true><<<(int)blocks, W * 32, smem, st>>>(c->d_tab, c->slot, hard, hard_dtype, B, ngroups, ok, sc.ptr);
    } else {
        LDPC_CUDA(cudaFuncSetAttribute(syndrome_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        syndrome_kernel<false><<<(int)blocks, W * 32, smem, st>>>(c->d_tab, 0, hard, hard_dtype, B, ngroups, ok, sc.ptr);
    }
    LDPC_CHECK_LAUNCH("syndrome_kernel");
    return LDPC_OK;
}

// ---- host launcher ----------------------------------------------------------------------
template <int kAlgo, int kMaxDc, int kMaxDv, bool kConst, bool kScratch>
inline int launch_exact_inst(const DecodeParams& p, int W, int grid, size_t smem, float* scratch, cudaStream_t st) {
    auto kern = decode_exact_kernel<kAlgo, kMaxDc, kMaxDv, kConst, kScratch>;
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, W * 32, smem, st>>>(p, scratch);
    LDPC_CHECK_LAUNCH("decode_exact_kernel");
    return LDPC_OK;
}

inline int launch_exact(const ldpc_code* c, int algo, DecodeParams p, cudaStream_t st) {
    if (c->maxdc > 32 || c->maxdv > 32)
        return fail(LDPC_ERR_UNSUPPORTED, "exact path: node degree above 32 (row %d, col %d)", c->maxdc, c->maxdv);
    const size_t per_warp = (size_t)(c->E + 2 * c->cols) * 32 * sizeof(float);
    int W = (int)(kMaxSmemPerBlock / per_warp);
    const long long ngroups = (p.B + c->G - 1) / c->G;
    ExactScratch sc;
    long long blocks;
    size_t smem = 0;
    if (W < (algo == LDPC_ALGO_BP ? 2 * LDPC_EXACT_WS_BELOW : LDPC_EXACT_WS_BELOW)) {
        // too few warps fit an SM's shared memory to hide the kernel's latencies: global workspace, 16 warps per SM (BG2 from
        // Z = 64 up as 32-circulants, large Z = 1 codes; sum-product, whose double-precision tanh / atanh chains are longer,
        // already at BG2 Z = 32 where five warps fit).  Measured, k cw/s, shared memory with what fits / workspace with 8 / with
        // 16 warps per SM: min-sum BG2 Z = 32 3288 / 2934 / 2794; sum-product 561 / 699 / 991; min-sum Z = 64 (2 warps fit)
        // 208 / 930 / 1138; Z = 128 (1 warp) 35 / 351 / 559; Z = 384 (none) - / 92 / 157
        if (int rc = sc.init(per_warp, ngroups, st, "exact path")) return rc;
        W = sc.W; blocks = sc.blocks;
    } else {
        if (W > 8) W = 8;
        if ((long long)W > ngroups) W = (int)ngroups;
        int per_sm = (int)(kMaxSmemPerBlock / (per_warp * W));
        per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
        blocks = (ngroups + W - 1) / W;
        if (blocks > (long long)kNumSMs * per_sm) blocks = (long long)kNumSMs * per_sm;
        smem = per_warp * W;
    }
    p.floats_per_warp = (int)(per_warp / sizeof(float));
    p.ngroups = ngroups;
    p.gtab = c->d_tab;
    p.slot = c->slot < 0 ? 0 : c->slot;
    const bool small = c->maxdc <= 10 && c->maxdv <= 24;
    const bool cst = c->slot >= 0;
#define LDPC_EXACT_CASE(A, DC, DV, CST) do { if (sc.ptr) return launch_exact_inst<A, DC, DV, CST, true>(p, W, (int)blocks, smem, sc.ptr, st); \
                                              return launch_exact_inst<A, DC, DV, CST, false>(p, W, (int)blocks, smem, nullptr, st); } while (0)
    if (algo == LDPC_ALGO_MINSUM) {
        if (small) { if (cst) LDPC_EXACT_CASE(LDPC_ALGO_MINSUM, 10, 24, true); else LDPC_EXACT_CASE(LDPC_ALGO_MINSUM, 10, 24, false); }
        else       { if (cst) LDPC_EXACT_CASE(LDPC_ALGO_MINSUM, 32, 32, true); else LDPC_EXACT_CASE(LDPC_ALGO_MINSUM, 32, 32, false); }
    } else {
        if (small) { if (cst) LDPC_EXACT_CASE(LDPC_ALGO_BP, 10, 24, true); else LDPC_EXACT_CASE(LDPC_ALGO_BP, 10, 24, false); }
        else       { if (cst) LDPC_EXACT_CASE(LDPC_ALGO_BP, 32, 32, true); else LDPC_EXACT_CASE(LDPC_ALGO_BP, 32, 32, false); }
    }
#undef LDPC_EXACT_CASE
    return LDPC_OK;                          // not reached: every case returns
}

}  // namespace ldpc
