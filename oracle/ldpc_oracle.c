/*
 * ldpc_oracle.c -- CPU restatement of the reference's flooding decoders and channel.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (ldpc-neuralnetwork-decoder_b200/)
 * may import, link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the timed CPU arm.
 *
 * What it restates (paths relative to /root/reference/ldpc_neural_decoder):
 *   oracle_decode, algo 0  MinSumScaledDecoder.decode      models/traditional_decoders.py:177-260
 *   oracle_decode, algo 1  BeliefPropagationDecoder.decode models/traditional_decoders.py:42-109
 *   validity masks          _check_valid_codeword           models/traditional_decoders.py:111-134,262-284
 *   oracle_awgn_llr         AWGNChannel.transmit            utils/channel.py:205-231 (noise source is
 *                           the engine's Philox generator, csrc/channel.cuh, not torch.randn)
 * The reference works on dense (B,M,N) tensors with Python loops; this file uses the Tanner
 * edge list of the QC code (check i*Z+r <-> variable j*Z+((r+s) mod Z), utils/ldpc_utils.py:121-123)
 * and performs the SAME fp32 operations in the SAME order per edge:
 *   - neighbours of a check are visited in ascending variable index, neighbours of a variable
 *     in ascending check index (the order of _precompute_indices, :26-40 / :161-175);
 *   - min-sum: signs = prod sign(v) (sign(0) = 0), min over the other edges, alpha*min in fp32,
 *     then signs*scaled (:207-232);
 *   - BP: prod of tanh(v/2) over the other edges, 2*atanh(prod), no clipping (:72-81);
 *     tanh/atanh are evaluated in double and rounded to fp32 (torch's CPU kernels differ from
 *     that by <= 1 ulp on ~0.4 % / 0.08 % of inputs -- measured);
 *   - variable update: llr + c2v of the other checks in ascending order (:235-244);
 *   - posterior: llr + every c2v in ascending order; bit = belief < 0 (:247-252).
 * order = 1 selects the engine's fast-path variable update (posterior minus own message) so
 * the specialised kernel can also be checked bit for bit.
 *
 * Pinned against the golden vectors produced by running the unmodified reference
 * (oracle/make_golden.py -> tests/golden/classic_*.npz, earlystop_*.npz): see
 * tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int rows, cols, Z, N, M, E;
    int *chk_ptr, *chk_var, *chk_edge; /* per check: variables ascending, edge ids               */
    int *var_ptr, *var_edge;           /* per variable: edge ids in ascending check order        */
    int *edge_var;                     /* edge id -> variable (edge ids are check-major)         */
} graph_t;

static int cmp_pair(const void* a, const void* b) {
    const int* x = (const int*)a;
    const int* y = (const int*)b;
    return x[0] != y[0] ? (x[0] > y[0]) - (x[0] < y[0]) : (x[1] > y[1]) - (x[1] < y[1]);
}

static graph_t* graph_build(const int16_t* shifts, int rows, int cols, int Z) {
    graph_t* g = (graph_t*)calloc(1, sizeof(graph_t));
    g->rows = rows; g->cols = cols; g->Z = Z; g->N = cols * Z; g->M = rows * Z;
    int be = 0;
    for (int i = 0; i < rows * cols; ++i) be += shifts[i] >= 0;
    g->E = be * Z;
    g->chk_ptr = (int*)calloc(g->M + 1, sizeof(int));
    g->chk_var = (int*)malloc(sizeof(int) * g->E);
    g->chk_edge = (int*)malloc(sizeof(int) * g->E);
    g->edge_var = (int*)malloc(sizeof(int) * g->E);
    g->var_ptr = (int*)calloc(g->N + 1, sizeof(int));
    g->var_edge = (int*)malloc(sizeof(int) * g->E);
    int e = 0;
    int* pairs = (int*)malloc(sizeof(int) * 2 * cols);
    for (int i = 0; i < rows; ++i)
        for (int r = 0; r < Z; ++r) {
            const int c = i * Z + r;
            int d = 0;
            for (int j = 0; j < cols; ++j) {
                const int s = shifts[i * cols + j];
                if (s < 0) continue;
                pairs[2 * d] = j * Z + (r + s) % Z;
                pairs[2 * d + 1] = 0;
                ++d;
            }
            qsort(pairs, d, 2 * sizeof(int), cmp_pair); /* ascending variable index */
            g->chk_ptr[c] = e;
            for (int k = 0; k < d; ++k) {
                g->chk_var[e] = pairs[2 * k];
                g->chk_edge[e] = e;
                g->edge_var[e] = pairs[2 * k];
                ++e;
            }
        }
    g->chk_ptr[g->M] = e;
    free(pairs);
    /* variable -> edges in ascending check order: edges are already check-major */
    for (int x = 0; x < g->E; ++x) g->var_ptr[g->edge_var[x] + 1]++;
    for (int v = 0; v < g->N; ++v) g->var_ptr[v + 1] += g->var_ptr[v];
    int* fill = (int*)calloc((size_t)g->N, sizeof(int));
    for (int x = 0; x < g->E; ++x) {
        const int v = g->edge_var[x];
        g->var_edge[g->var_ptr[v] + fill[v]++] = x;
    }
    free(fill);
    return g;
}

static void graph_free(graph_t* g) {
    free(g->chk_ptr); free(g->chk_var); free(g->chk_edge); free(g->edge_var); free(g->var_ptr); free(g->var_edge);
    free(g);
}

/* torch.sign: (0 < x) - (x < 0), hence sign(NaN) = 0 (pinned by tests/golden/nonfinite_z4_b24.npz) */
static inline float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float tanh_ref(float x) { return (float)tanh((double)x); }
static inline float atanh_ref(float x) { return (float)atanh((double)x); }

static int all_checks_ok(const graph_t* g, const float* belief) {
    for (int c = 0; c < g->M; ++c) {
        int par = 0;
        for (int k = g->chk_ptr[c]; k < g->chk_ptr[c + 1]; ++k) par ^= belief[g->chk_var[k]] < 0.0f;
        if (par) return 0;
    }
    return 1;
}

/* one codeword, `iters` flooding iterations; v2c/c2v are per-edge scratch of size E */
static void decode_one(const graph_t* g, int algo, int order, const float* llr, int iters, float alpha, float* v2c,
                       float* c2v, float* belief, uint64_t* valid_mask, int mask_words, int stop_when_valid,
                       int* iters_done) {
    const volatile float alpha_f = alpha;
    for (int x = 0; x < g->E; ++x) v2c[x] = llr[g->edge_var[x]];
    memcpy(belief, llr, sizeof(float) * g->N);
    if (valid_mask) memset(valid_mask, 0, sizeof(uint64_t) * mask_words);
    int done_at = iters;
    for (int it = 0; it < iters; ++it) {
        /* check-node update */
        for (int c = 0; c < g->M; ++c) {
            const int k0 = g->chk_ptr[c], k1 = g->chk_ptr[c + 1];
            for (int k = k0; k < k1; ++k) {
                if (algo == 0) {
                    float signs = 1.0f, mn = INFINITY;
                    for (int q = k0; q < k1; ++q) {
                        if (q == k) continue;
                        signs = signs * sgnf(v2c[q]);
                        const float mag = fabsf(v2c[q]);
                        if (mag < mn) mn = mag;
                    }
                    const float scaled = alpha_f * mn;
                    c2v[k] = signs * scaled;
                } else {
                    float prod = 1.0f;
                    for (int q = k0; q < k1; ++q) {
                        if (q == k) continue;
                        prod = prod * tanh_ref(v2c[q] / 2.0f);
                    }
                    c2v[k] = 2.0f * atanh_ref(prod);
                }
            }
        }
        /* posterior, ascending check order */
        for (int v = 0; v < g->N; ++v) {
            float b = llr[v];
            for (int k = g->var_ptr[v]; k < g->var_ptr[v + 1]; ++k) b = b + c2v[g->var_edge[k]];
            belief[v] = b;
        }
        /* variable-node update */
        for (int v = 0; v < g->N; ++v) {
            const int k0 = g->var_ptr[v], k1 = g->var_ptr[v + 1];
            for (int k = k0; k < k1; ++k) {
                float s;
                if (order == 1 && k1 - k0 > 1) {
                    s = belief[v] - c2v[g->var_edge[k]]; /* engine fast path: total minus self */
                } else {
                    s = llr[v];
                    for (int q = k0; q < k1; ++q)
                        if (q != k) s = s + c2v[g->var_edge[q]];
                }
                v2c[g->var_edge[k]] = s;
            }
        }
        if (valid_mask || stop_when_valid) {
            const int ok = all_checks_ok(g, belief);
            if (ok && valid_mask && (it >> 6) < mask_words) valid_mask[it >> 6] |= 1ull << (it & 63);
            if (ok && stop_when_valid) { done_at = it + 1; break; }
        }
    }
    if (iters_done) *iters_done = done_at;
}

/* Batch decode.  beliefs [B,N] (may be NULL), hard [B,N] uint8 (may be NULL), valid_mask
 * [B,mask_words] (may be NULL), iters_done [B] (may be NULL).  stop_when_valid = per-codeword
 * early exit.  threads <= 0: all OpenMP threads.  Returns 0, or -1 on bad arguments. */
int oracle_decode(const int16_t* shifts, int rows, int cols, int Z, int algo, int order, const float* llr, int64_t B,
                  int iters, float alpha, float* beliefs, uint8_t* hard, uint64_t* valid_mask, int mask_words,
                  int stop_when_valid, int32_t* iters_done, int threads) {
    if (!shifts || !llr || iters < 1 || Z < 1 || B < 0) return -1;
    graph_t* g = graph_build(shifts, rows, cols, Z);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        float* v2c = (float*)malloc(sizeof(float) * g->E);
        float* c2v = (float*)malloc(sizeof(float) * g->E);
        float* bel = (float*)malloc(sizeof(float) * g->N);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            int done = 0;
            decode_one(g, algo, order, llr + b * g->N, iters, alpha, v2c, c2v, bel,
                       valid_mask ? valid_mask + b * mask_words : NULL, mask_words, stop_when_valid, &done);
            if (beliefs) memcpy(beliefs + b * g->N, bel, sizeof(float) * g->N);
            if (hard)
                for (int v = 0; v < g->N; ++v) hard[b * g->N + v] = bel[v] < 0.0f;
            if (iters_done) iters_done[b] = done;
        }
        free(v2c); free(c2v); free(bel);
    }
    graph_free(g);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- channel: Philox4x32-10 + Box-Muller, the generator of csrc/channel.cuh ------------- */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float u01(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

int oracle_awgn_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, uint64_t seed, uint64_t first_frame,
                    float* out) {
    if (!out || B < 0 || N <= 0) return -1;
    const double snr_linear = pow(10.0, (double)snr_db / 10.0);
    const double sigma_d = 1.0 / sqrt(snr_linear);
    const volatile float sigma = (float)sigma_d, var = (float)(sigma_d * sigma_d);
    const int64_t nblk = ((N + 127) >> 7) << 5;
    for (int64_t b = 0; b < B; ++b) {
        const uint64_t frame = first_frame + (uint64_t)b;
        for (int64_t blk = 0; blk < nblk; ++blk) {
            uint32_t x[4];
            philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)blk, 0u, (uint32_t)seed,
                          (uint32_t)(seed >> 32), x);
            const float r0 = sqrtf(-2.0f * logf(u01(x[0]))), r1 = sqrtf(-2.0f * logf(u01(x[2])));
            const float a0 = 6.2831853071795865f * u01(x[1]), a1 = 6.2831853071795865f * u01(x[3]);
            const float z[4] = {r0 * cosf(a0), r0 * sinf(a0), r1 * cosf(a1), r1 * sinf(a1)};
            for (int comp = 0; comp < 4; ++comp) {
                const int64_t n = ((blk >> 5) << 7) + ((int64_t)comp << 5) + (blk & 31);
                if (n >= N) continue;
                const float s = bits ? 1.0f - 2.0f * (float)bits[b * N + n] : 1.0f;
                const float noise = z[comp] * sigma;
                const float received = s + noise;
                const float twice = 2.0f * received;
                out[b * N + n] = twice / var;
            }
        }
    }
    return 0;
}

/* qpsk_modulate -> awgn_channel -> qpsk_demodulate, utils/channel.py:39, 75-82, 120-138 (noise: the engine's Philox) */
int oracle_qpsk_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, int true_llr, uint64_t seed,
                    uint64_t first_frame, float* out) {
    if (!out || B < 0 || N <= 0) return -1;
    const double snr_linear = pow(10.0, (double)snr_db / 10.0);
    const double noise_power = 1.0 / snr_linear;
    const volatile float amp = (float)(1.0 / sqrt(2.0)), sigma = (float)sqrt(noise_power / 2.0),
                         var = true_llr ? (float)(noise_power / sqrt(2.0)) : (float)noise_power;
    const int64_t nblk = ((N + 127) >> 7) << 5;
    for (int64_t b = 0; b < B; ++b) {
        const uint64_t frame = first_frame + (uint64_t)b;
        for (int64_t blk = 0; blk < nblk; ++blk) {
            uint32_t x[4];
            philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), (uint32_t)blk, 0u, (uint32_t)seed,
                          (uint32_t)(seed >> 32), x);
            const float r0 = sqrtf(-2.0f * logf(u01(x[0]))), r1 = sqrtf(-2.0f * logf(u01(x[2])));
            const float a0 = 6.2831853071795865f * u01(x[1]), a1 = 6.2831853071795865f * u01(x[3]);
            const float z[4] = {r0 * cosf(a0), r0 * sinf(a0), r1 * cosf(a1), r1 * sinf(a1)};
            for (int comp = 0; comp < 4; ++comp) {
                const int64_t n = ((blk >> 5) << 7) + ((int64_t)comp << 5) + (blk & 31);
                if (n >= N) continue;
                const float s = bits && bits[b * N + n] ? -amp : amp;
                const float noise = z[comp] * sigma;
                const float received = s + noise;
                const float twice = 2.0f * received;
                out[b * N + n] = twice / var;
            }
        }
    }
    return 0;
}
