// ldpc_b200.cu -- the C ABI of include/ldpc_b200.h.  Host-side argument checking, table
// upload and kernel dispatch only; the kernels live in the .cuh files next to this one.
// Built for sm_100a only (see __graft_entry__.build / Makefile); there is no CPU fallback.
#include "common.cuh"
#include "tables.cuh"
#include "decode_exact.cuh"
#include "decode_fast.cuh"
#include "channel_kernels.cuh"
#include "layers.cuh"
#include "neural.cuh"
#include "neural_qc.cuh"
#include "gnn.cuh"
#include "gnn_bwd.cuh"
#include <cuda_fp16.h>
#include "encode.cuh"
#include "rate_match.cuh"
#include "gnn_tc.cuh"
#include "gnn_tc_pipe.cuh"
#include "gnn_node_pipe.cuh"
#include "gnn_bwd_tc.cuh"
#include <cstdlib>

#include <cstring>
#include <new>

using namespace ldpc;

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_decode_args(const ldpc_code_t* code, const float* llr, int64_t B, int iters, int stop_mode, int path,
                      int hard_dtype, const uint64_t* valid_mask, int mask_words) {
    if (!code) return fail(LDPC_ERR_INVALID, "decode: null code handle");
    if (B < 0) return fail(LDPC_ERR_INVALID, "decode: negative batch %lld", (long long)B);
    if (!llr && B > 0) return fail(LDPC_ERR_INVALID, "decode: null llr pointer");
    if (iters < 1) return fail(LDPC_ERR_INVALID, "decode: iters=%d, need >= 1", iters);
    if (stop_mode != LDPC_STOP_FIXED && stop_mode != LDPC_STOP_PER_CODEWORD)
        return fail(LDPC_ERR_INVALID, "decode: unknown stop_mode %d", stop_mode);
    if (path < LDPC_PATH_AUTO || path > LDPC_PATH_FAST) return fail(LDPC_ERR_INVALID, "decode: unknown path %d", path);
    if (hard_dtype < LDPC_HARD_F32 || hard_dtype > LDPC_HARD_PACKED)
        return fail(LDPC_ERR_INVALID, "decode: unknown hard_dtype %d", hard_dtype);
    if (valid_mask && mask_words < (iters + 63) / 64)
        return fail(LDPC_ERR_INVALID, "decode: mask_words=%d too small for %d iterations", mask_words, iters);
    return LDPC_OK;
}

int decode_common(const ldpc_code_t* code, int algo, const float* llr, int64_t B, int iters, float alpha, int stop_mode,
                  int path, float* soft_out, void* hard_out, int hard_dtype, uint8_t* syndrome_ok, int32_t* iters_out,
                  uint64_t* valid_mask, int mask_words, void* stream) {
    int rc = check_decode_args(code, llr, B, iters, stop_mode, path, hard_dtype, valid_mask, mask_words);
    if (rc) return rc;
    if (B == 0) return LDPC_OK;
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "decode: cannot select device %d", code->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (valid_mask) LDPC_CUDA(cudaMemsetAsync(valid_mask, 0, sizeof(uint64_t) * (size_t)B * mask_words, st));
    DecodeParams p{};
    p.llr = llr; p.B = B; p.iters = iters; p.alpha = alpha; p.stop_mode = stop_mode;
    p.soft_out = soft_out; p.hard_out = hard_out; p.hard_dtype = hard_dtype; p.syndrome_ok = syndrome_ok;
    p.iters_out = iters_out; p.valid_mask = (unsigned long long*)valid_mask; p.mask_words = mask_words;
    const bool fast_ok = fast_path_supports(code, algo, stop_mode, valid_mask != nullptr, soft_out != nullptr);
    if (path == LDPC_PATH_FAST && !fast_ok)
        return fail(LDPC_ERR_UNSUPPORTED, "decode: no specialised kernel for this code/algorithm/stop mode");
    // AUTO: the specialised kernel where one exists.  Min-sum: hard decisions identical to the reference-order kernel on
    // the bench's 2^20 frames, soft outputs to rounding.  BP (10 M vs 0.34 M codewords/s): 0 hard-bit mismatches and 22
    // inf/NaN class mismatches in 1.7e9 beliefs on the same frames (bench.py --workload bp, `parity`); LDPC_PATH_EXACT
    // keeps the reference's operation order and its once-rounded tanh/atanh.
    if (fast_ok && path != LDPC_PATH_EXACT) return launch_fast(code, algo, p, st);
    return launch_exact(code, algo, p, st);
}

}  // namespace

extern "C" {

int ldpc_abi_version(void) { return LDPC_B200_ABI_VERSION; }
const char* ldpc_last_error(void) { return err_buf(); }
uint64_t ldpc_launch_count(void) { return launch_counter().load(); }

int ldpc_code_create(const int16_t* shifts, int rows, int cols, int Z, int device, ldpc_code_t** out) {
    if (!out) return fail(LDPC_ERR_INVALID, "code_create: null out pointer");
    *out = nullptr;
    if (device < 0 || device >= kMaxDevices) return fail(LDPC_ERR_INVALID, "code_create: device %d out of range", device);
    ldpc_code* c = new (std::nothrow) ldpc_code();
    if (!c) return fail(LDPC_ERR_NOMEM, "code_create: out of host memory");
    int rc = build_table(shifts, rows, cols, Z, c);
    if (rc) { delete c; return rc; }
    c->device = device;
    DeviceGuard g(device);
    if (!g.ok) { delete c; return fail(LDPC_ERR_CUDA, "code_create: cannot select CUDA device %d (no GPU?)", device); }
    cudaError_t e = cudaMalloc(&c->d_tab, sizeof(uint32_t) * c->tab_words);
    if (e != cudaSuccess) { delete c; return fail(LDPC_ERR_CUDA, "code_create: cudaMalloc: %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(c->d_tab, c->h_tab.data(), sizeof(uint32_t) * c->tab_words, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(c->d_tab); delete c; return fail(LDPC_ERR_CUDA, "code_create: upload: %s", cudaGetErrorString(e)); }
    if (c->tab_words <= kSlotWords) {
        std::lock_guard<std::mutex> lk(slot_mutex());
        for (int s = 0; s < kNumSlots; ++s)
            if (!slot_used()[device][s]) {
                e = cudaMemcpyToSymbol(c_tab, c->h_tab.data(), sizeof(uint32_t) * c->tab_words,
                                       sizeof(uint32_t) * (size_t)s * kSlotWords, cudaMemcpyHostToDevice);
                if (e == cudaSuccess) { slot_used()[device][s] = true; c->slot = s; }
                break;
            }
    }
    c->fast_kind = detect_fast_kind(c);
    *out = c;
    return LDPC_OK;
}

int ldpc_code_destroy(ldpc_code_t* code) {
    if (!code) return LDPC_OK;
    {
        DeviceGuard g(code->device);
        if (code->d_tab) cudaFree(code->d_tab);
        for (int s = 0; s < HostStage::kStages; ++s) {
            if (code->stage.st[s]) { cudaStreamSynchronize(code->stage.st[s]); cudaStreamDestroy(code->stage.st[s]); }
            cudaFree(code->stage.d_llr[s]); cudaFree(code->stage.d_hard[s]); cudaFree(code->stage.d_soft[s]); cudaFree(code->stage.d_raw[s]);
        }
    }
    if (code->slot >= 0) {
        std::lock_guard<std::mutex> lk(slot_mutex());
        slot_used()[code->device][code->slot] = false;
    }
    delete code;
    return LDPC_OK;
}

int ldpc_code_info(const ldpc_code_t* c, int32_t info[8]) {
    if (!c || !info) return fail(LDPC_ERR_INVALID, "code_info: null argument");
    info[0] = c->rows; info[1] = c->cols; info[2] = c->Z; info[3] = c->E;
    info[4] = c->N; info[5] = c->M; info[6] = c->maxdc; info[7] = c->maxdv;
    return LDPC_OK;
}

int ldpc_code_has_fast_path(const ldpc_code_t* code, int algo) {
    if (!code) return 0;
    return fast_path_supports(code, algo, LDPC_STOP_FIXED, false, false) ? 1 : 0;
}

int ldpc_minsum_decode(const ldpc_code_t* code, const float* llr, int64_t B, int iters, float alpha, int stop_mode,
                       int path, float* soft_out, void* hard_out, int hard_dtype, uint8_t* syndrome_ok,
                       int32_t* iters_out, uint64_t* valid_mask, int mask_words, void* stream) {
    return decode_common(code, LDPC_ALGO_MINSUM, llr, B, iters, alpha, stop_mode, path, soft_out, hard_out, hard_dtype,
                         syndrome_ok, iters_out, valid_mask, mask_words, stream);
}

int ldpc_bp_decode(const ldpc_code_t* code, const float* llr, int64_t B, int iters, int stop_mode, int path,
                   float* soft_out, void* hard_out, int hard_dtype, uint8_t* syndrome_ok, int32_t* iters_out,
                   uint64_t* valid_mask, int mask_words, void* stream) {
    return decode_common(code, LDPC_ALGO_BP, llr, B, iters, 1.0f, stop_mode, path, soft_out, hard_out, hard_dtype,
                         syndrome_ok, iters_out, valid_mask, mask_words, stream);
}

int ldpc_syndrome_check(const ldpc_code_t* code, const void* hard, int hard_dtype, int64_t B, uint8_t* syndrome_ok,
                        void* stream) {
    if (!code) return fail(LDPC_ERR_INVALID, "syndrome_check: null code handle");
    if (B < 0) return fail(LDPC_ERR_INVALID, "syndrome_check: negative batch");
    if (B == 0) return LDPC_OK;
    if (!hard || !syndrome_ok) return fail(LDPC_ERR_INVALID, "syndrome_check: null buffer");
    if (hard_dtype < LDPC_HARD_F32 || hard_dtype > LDPC_HARD_PACKED) return fail(LDPC_ERR_INVALID, "syndrome_check: unknown hard_dtype");
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "syndrome_check: cannot select device %d", code->device);
    return launch_syndrome(code, hard, hard_dtype, B, syndrome_ok, (cudaStream_t)stream);
}

int ldpc_encode(const ldpc_code_t* code, const uint8_t* info, int64_t B, const int32_t* plan, int64_t plan_len,
                const uint32_t* binv, uint8_t* codeword, void* stream) {
    if (!code || !plan || !binv) return fail(LDPC_ERR_INVALID, "encode: null code handle or encoder tables");
    if (B < 0) return fail(LDPC_ERR_INVALID, "encode: negative batch");
    if (B == 0) return LDPC_OK;
    if (!info || !codeword) return fail(LDPC_ERR_INVALID, "encode: null buffer");
    if (plan_len < 3 + code->rows) return fail(LDPC_ERR_INVALID, "encode: plan of %lld words is too short", (long long)plan_len);
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "encode: cannot select device %d", code->device);
    // words per row of B^-1: the plan carries it at [2]; it is needed on the host for the shared-memory size
    int hdr[3];
    LDPC_CUDA(cudaMemcpyAsync(hdr, plan, sizeof(hdr), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LDPC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] != (hdr[0] * code->Z + 31) / 32 || plan_len != 3 + 2 * hdr[0] + code->rows)
        return fail(LDPC_ERR_INVALID, "encode: inconsistent plan (g %d, kb %d, words %d, length %lld)", hdr[0], hdr[1], hdr[2],
                    (long long)plan_len);
    return launch_encode(code, info, B, plan, hdr[2], binv, codeword, (cudaStream_t)stream);
}

int ldpc_rate_match(const uint8_t* codeword, const int32_t* sel, int64_t B, int64_t N, int64_t E, uint8_t* out, void* stream) {
    if (!codeword || !sel || !out) return fail(LDPC_ERR_INVALID, "rate_match: null argument");
    if (B < 0 || N <= 0 || E <= 0 || N > 0x7fffffff || E > 0x7fffffff) return fail(LDPC_ERR_INVALID, "rate_match: bad shape");
    if (B == 0) return LDPC_OK;
    const long long blocks = (B * E + 255) / 256;
    rate_match_kernel<<<(int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        codeword, sel, (long long)B, (int)N, (int)E, out);
    LDPC_CHECK_LAUNCH("rate_match_kernel");
    return LDPC_OK;
}

int ldpc_rate_recover(const float* rx_llr, const int32_t* inv_ptr, const int32_t* inv_idx, const float* base, int64_t B,
                      int64_t N, int64_t E, float* llr_out, void* stream) {
    if (!rx_llr || !inv_ptr || !inv_idx || !base || !llr_out) return fail(LDPC_ERR_INVALID, "rate_recover: null argument");
    if (B < 0 || N <= 0 || E <= 0 || N > 0x7fffffff || E > 0x7fffffff) return fail(LDPC_ERR_INVALID, "rate_recover: bad shape");
    if (B == 0) return LDPC_OK;
    const long long blocks = (B * N + 255) / 256;
    rate_recover_kernel<<<(int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        rx_llr, inv_ptr, inv_idx, base, (long long)B, (int)N, (int)E, llr_out);
    LDPC_CHECK_LAUNCH("rate_recover_kernel");
    return LDPC_OK;
}

int ldpc_decode_host_q(const ldpc_code_t* code, int algo, const void* llr_host, int llr_format, float llr_scale,
                       int64_t B, int iters, float alpha, int path, float* soft_host, void* hard_host,
                       int hard_dtype, int64_t chunk) {
    if (!code) return fail(LDPC_ERR_INVALID, "decode_host: null code handle");
    if (B < 0) return fail(LDPC_ERR_INVALID, "decode_host: negative batch");
    if (B == 0) return LDPC_OK;
    if (!llr_host || !hard_host) return fail(LDPC_ERR_INVALID, "decode_host: null host buffer");
    if (algo != LDPC_ALGO_MINSUM && algo != LDPC_ALGO_BP) return fail(LDPC_ERR_INVALID, "decode_host: unknown algo %d", algo);
    if (hard_dtype < LDPC_HARD_F32 || hard_dtype > LDPC_HARD_PACKED) return fail(LDPC_ERR_INVALID, "decode_host: unknown hard_dtype");
    if (llr_format < LDPC_LLR_F32 || llr_format > LDPC_LLR_I8) return fail(LDPC_ERR_INVALID, "decode_host: unknown llr_format %d", llr_format);
    const size_t raw_elem = llr_format == LDPC_LLR_I8 ? 1 : llr_format == LDPC_LLR_F16 ? 2 : 0;   // 0: fp32 goes straight to d_llr
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "decode_host: cannot select device %d", code->device);
    const int N = code->N;
    if (chunk <= 0) chunk = 1 << 15;
    if (chunk > B) chunk = B;
    const size_t hard_row = hard_dtype == LDPC_HARD_F32 ? sizeof(float) * N
                          : hard_dtype == LDPC_HARD_U8 ? (size_t)N : sizeof(uint32_t) * ((N + 31) / 32);
    // staging buffers and streams live in the handle and are reused by later calls
    HostStage& hs = code->stage;
    std::lock_guard<std::mutex> lk(hs.mu);
    const size_t need_llr = sizeof(float) * (size_t)chunk * N, need_hard = hard_row * (size_t)chunk;
    const size_t need_soft = soft_host ? need_llr : 0, need_raw = raw_elem * (size_t)chunk * N;
    for (int s = 0; s < HostStage::kStages; ++s) {
        if (!hs.st[s]) LDPC_CUDA(cudaStreamCreateWithFlags(&hs.st[s], cudaStreamNonBlocking));
        if (hs.cap_llr[s] < need_llr) {
            LDPC_CUDA(cudaStreamSynchronize(hs.st[s]));
            cudaFree(hs.d_llr[s]); hs.d_llr[s] = nullptr; hs.cap_llr[s] = 0;
            LDPC_CUDA(cudaMalloc(&hs.d_llr[s], need_llr)); hs.cap_llr[s] = need_llr;
        }
        if (hs.cap_raw[s] < need_raw) {
            LDPC_CUDA(cudaStreamSynchronize(hs.st[s]));
            cudaFree(hs.d_raw[s]); hs.d_raw[s] = nullptr; hs.cap_raw[s] = 0;
            LDPC_CUDA(cudaMalloc(&hs.d_raw[s], need_raw)); hs.cap_raw[s] = need_raw;
        }
        if (hs.cap_hard[s] < need_hard) {
            LDPC_CUDA(cudaStreamSynchronize(hs.st[s]));
            cudaFree(hs.d_hard[s]); hs.d_hard[s] = nullptr; hs.cap_hard[s] = 0;
            LDPC_CUDA(cudaMalloc(&hs.d_hard[s], need_hard)); hs.cap_hard[s] = need_hard;
        }
        if (hs.cap_soft[s] < need_soft) {
            LDPC_CUDA(cudaStreamSynchronize(hs.st[s]));
            cudaFree(hs.d_soft[s]); hs.d_soft[s] = nullptr; hs.cap_soft[s] = 0;
            LDPC_CUDA(cudaMalloc(&hs.d_soft[s], need_soft)); hs.cap_soft[s] = need_soft;
        }
    }
    // error inside the chunk loop: async D2H copies into the CALLER's buffers may still be in flight on the staging
    // streams, so settle all of them before the error code goes back
    auto settle = [&hs]() { for (int s = 0; s < HostStage::kStages; ++s) if (hs.st[s]) cudaStreamSynchronize(hs.st[s]); };
#define LDPC_CUDA_SETTLE(expr)                                                                                     \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess) {                                                                                   \
            settle();                                                                                              \
            return fail(LDPC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                                          \
    } while (0)
    int64_t done = 0;
    for (int it = 0; done < B; ++it, done += chunk) {
        const int s = it % HostStage::kStages;
        const int64_t b = (B - done) < chunk ? (B - done) : chunk;
        // same-stream ordering makes reuse of stage s safe: its previous D2H precedes this H2D
        if (raw_elem == 0) {
            LDPC_CUDA_SETTLE(cudaMemcpyAsync(hs.d_llr[s], (const float*)llr_host + (size_t)done * N, sizeof(float) * (size_t)b * N,
                                      cudaMemcpyHostToDevice, hs.st[s]));
        } else {
            const long long n = (long long)b * N;
            LDPC_CUDA_SETTLE(cudaMemcpyAsync(hs.d_raw[s], (const char*)llr_host + raw_elem * (size_t)done * N, raw_elem * (size_t)n,
                                      cudaMemcpyHostToDevice, hs.st[s]));
            const int grid = (int)((n + 1023) / 1024 < (long long)kNumSMs * 8 ? (n + 1023) / 1024 : (long long)kNumSMs * 8);
            if (llr_format == LDPC_LLR_I8)
                llr_dequant_kernel<int8_t><<<grid, 256, 0, hs.st[s]>>>((const int8_t*)hs.d_raw[s], llr_scale, n, (float*)hs.d_llr[s]);
            else
                llr_dequant_kernel<__half><<<grid, 256, 0, hs.st[s]>>>((const __half*)hs.d_raw[s], llr_scale, n, (float*)hs.d_llr[s]);
            LDPC_COUNT_LAUNCH();
            LDPC_CUDA_SETTLE(cudaGetLastError());
        }
        int rc = decode_common(code, algo, (const float*)hs.d_llr[s], b, iters, alpha, LDPC_STOP_FIXED, path,
                               soft_host ? (float*)hs.d_soft[s] : nullptr, hs.d_hard[s], hard_dtype, nullptr, nullptr,
                               nullptr, 0, hs.st[s]);
        if (rc) { settle(); return rc; }
        LDPC_CUDA_SETTLE(cudaMemcpyAsync((char*)hard_host + hard_row * (size_t)done, hs.d_hard[s], hard_row * (size_t)b,
                                  cudaMemcpyDeviceToHost, hs.st[s]));
        if (soft_host)
            LDPC_CUDA_SETTLE(cudaMemcpyAsync(soft_host + (size_t)done * N, hs.d_soft[s], sizeof(float) * (size_t)b * N,
                                      cudaMemcpyDeviceToHost, hs.st[s]));
    }
    for (int s = 0; s < HostStage::kStages; ++s) LDPC_CUDA(cudaStreamSynchronize(hs.st[s]));
#undef LDPC_CUDA_SETTLE
    return LDPC_OK;
}

int ldpc_decode_host(const ldpc_code_t* code, int algo, const float* llr_host, int64_t B, int iters, float alpha,
                     int path, float* soft_host, void* hard_host, int hard_dtype, int64_t chunk) {
    return ldpc_decode_host_q(code, algo, llr_host, LDPC_LLR_F32, 1.0f, B, iters, alpha, path, soft_host, hard_host, hard_dtype, chunk);
}

// ---- channel + metrics ------------------------------------------------------------------
static int make_gen(float snr_db, uint64_t seed, uint64_t first_frame, GenParams* g) {
    // utils/channel.py:219-222: snr_linear = 10**(snr_db/10); noise_std = 1/np.sqrt(snr_linear), float64
    const double snr_linear = pow(10.0, (double)snr_db / 10.0);
    const double sigma = 1.0 / sqrt(snr_linear);
    g->enabled = 1;
    g->sigma = (float)sigma;
    g->var = (float)(sigma * sigma);
    g->amp = 1.0f;
    g->seed = seed;
    g->first_frame = first_frame;
    return LDPC_OK;
}

int ldpc_awgn_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, uint64_t seed, uint64_t first_frame,
                  float* llr_out, void* stream) {
    if (!llr_out) return fail(LDPC_ERR_INVALID, "awgn_llr: null output");
    if (B < 0 || N <= 0) return fail(LDPC_ERR_INVALID, "awgn_llr: bad shape [%lld,%lld]", (long long)B, (long long)N);
    if (B == 0) return LDPC_OK;
    GenParams g;
    make_gen(snr_db, seed, first_frame, &g);
    const long long blocks_needed = (B * ((((long long)N + 127) >> 7) << 5) + 255) / 256;
    const int grid = (int)(blocks_needed < kNumSMs * 8 ? blocks_needed : kNumSMs * 8);
    awgn_llr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits, B, N, g, llr_out);
    LDPC_CHECK_LAUNCH("awgn_llr_kernel");
    return LDPC_OK;
}

int ldpc_qpsk_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, int true_llr, uint64_t seed,
                  uint64_t first_frame, float* llr_out, void* stream) {
    if (!llr_out) return fail(LDPC_ERR_INVALID, "qpsk_llr: null output");
    if (B < 0 || N <= 0) return fail(LDPC_ERR_INVALID, "qpsk_llr: bad shape [%lld,%lld]", (long long)B, (long long)N);
    if (B == 0) return LDPC_OK;
    // utils/channel.py:39 (symbol components +-1/sqrt(2)), :75-82 (noise_power = 1/snr_linear, each component
    // N(0, noise_power/2)), :120-138 (llr = 2*r/noise_var with noise_var = 1/snr_linear): all scalars are Python
    // float64 rounded to fp32 where they meet a tensor
    const double snr_linear = pow(10.0, (double)snr_db / 10.0);
    const double noise_power = 1.0 / snr_linear;
    GenParams g;
    g.enabled = 1;
    g.amp = (float)(1.0 / sqrt(2.0));
    g.sigma = (float)sqrt(noise_power / 2.0);
    // the reference divides by the TOTAL noise variance although each component carries half of it and the symbol
    // amplitude is 1/sqrt(2): its LLRs are 1/sqrt(2) of the true ones.  true_llr = 1 divides by noise_power/sqrt(2).
    g.var = true_llr ? (float)(noise_power / sqrt(2.0)) : (float)noise_power;
    g.seed = seed;
    g.first_frame = first_frame;
    const long long blocks_needed = (B * ((((long long)N + 127) >> 7) << 5) + 255) / 256;
    const int grid = (int)(blocks_needed < kNumSMs * 8 ? blocks_needed : kNumSMs * 8);
    awgn_llr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits, B, N, g, llr_out);
    LDPC_CHECK_LAUNCH("awgn_llr_kernel(qpsk)");
    return LDPC_OK;
}

int ldpc_nonfinite_flag(const float* x, int64_t n, int32_t* flag, void* stream) {
    if (!x || !flag) return fail(LDPC_ERR_INVALID, "nonfinite_flag: null argument");
    if (n < 0) return fail(LDPC_ERR_INVALID, "nonfinite_flag: negative length");
    if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return fail(LDPC_ERR_INVALID, "nonfinite_flag: x must be 16-byte aligned");
    if (n == 0) return LDPC_OK;
    const long long blocks = ((n >> 2) + 255) / 256;
    nonfinite_flag_kernel<<<(int)(blocks < 1 ? 1 : (blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8)), 256, 0, (cudaStream_t)stream>>>(
        x, (long long)n, flag);
    LDPC_CHECK_LAUNCH("nonfinite_flag_kernel");
    return LDPC_OK;
}

int ldpc_count_errors(const void* hard, int hard_dtype, const uint8_t* tx, int64_t B, int64_t N, uint64_t* counters,
                      void* stream) {
    if (!hard || !counters) return fail(LDPC_ERR_INVALID, "count_errors: null argument");
    if (hard_dtype < LDPC_HARD_F32 || hard_dtype > LDPC_HARD_PACKED) return fail(LDPC_ERR_INVALID, "count_errors: unknown hard_dtype");
    if (B < 0 || N <= 0) return fail(LDPC_ERR_INVALID, "count_errors: bad shape");
    if (B == 0) return LDPC_OK;
    const long long blocks_needed = (B + 7) / 8;
    const int grid = (int)(blocks_needed < kNumSMs * 8 ? blocks_needed : kNumSMs * 8);
    count_errors_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hard, hard_dtype, tx, B, N, (unsigned long long*)counters);
    LDPC_CHECK_LAUNCH("count_errors_kernel");
    return LDPC_OK;
}

int ldpc_sim_fer(const ldpc_code_t* code, int algo, int iters, float alpha, float snr_db, uint64_t seed,
                 uint64_t first_frame, uint64_t n_frames, uint64_t* counters, void* stream) {
    if (!code || !counters) return fail(LDPC_ERR_INVALID, "sim_fer: null argument");
    if (algo != LDPC_ALGO_MINSUM && algo != LDPC_ALGO_BP) return fail(LDPC_ERR_INVALID, "sim_fer: unknown algo %d", algo);
    if (iters < 1) return fail(LDPC_ERR_INVALID, "sim_fer: iters=%d", iters);
    if (n_frames == 0) return LDPC_OK;
    if (n_frames > (1ull << 40)) return fail(LDPC_ERR_INVALID, "sim_fer: more than 2^40 frames in one call");
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "sim_fer: cannot select device %d", code->device);
    DecodeParams p{};
    p.B = (long long)n_frames; p.iters = iters; p.alpha = alpha; p.stop_mode = LDPC_STOP_FIXED;
    p.counters = (unsigned long long*)counters;
    make_gen(snr_db, seed, first_frame, &p.gen);
    if (fast_path_supports(code, algo, LDPC_STOP_FIXED, false, false)) return launch_fast(code, algo, p, (cudaStream_t)stream);
    return launch_exact(code, algo, p, (cudaStream_t)stream);
}

// ---- edge-space layers ---------------------------------------------------------------------
#define LDPC_LAYER_DISPATCH(KERNEL, IDX, ...)                                                            \
    do {                                                                                            \
        const int rows = layer_rows_per_cta(E);                                                     \
        const int grid = layer_grid(rows ? (B + rows - 1) / rows : B, 1);                           \
        const size_t smem = (size_t)rows * E * sizeof(float);                                       \
        switch (rows) {                                                                             \
            case 0: KERNEL<1, false, IDX><<<grid, kLayerThreads, 0, st>>>(__VA_ARGS__); break;           \
            case 1: LDPC_CUDA(cudaFuncSetAttribute(KERNEL<1, true, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                    KERNEL<1, true, IDX><<<grid, kLayerThreads, smem, st>>>(__VA_ARGS__); break;         \
            case 2: case 3: { const size_t sm2 = (size_t)2 * E * sizeof(float);                     \
                    LDPC_CUDA(cudaFuncSetAttribute(KERNEL<2, true, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2)); \
                    KERNEL<2, true, IDX><<<layer_grid((B + 1) / 2, 1), kLayerThreads, sm2, st>>>(__VA_ARGS__); break; } \
            case 4: case 5: case 6: case 7: { const size_t sm4 = (size_t)4 * E * sizeof(float);     \
                    LDPC_CUDA(cudaFuncSetAttribute(KERNEL<4, true, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4)); \
                    KERNEL<4, true, IDX><<<layer_grid((B + 3) / 4, 1), kLayerThreads, sm4, st>>>(__VA_ARGS__); break; } \
            default: LDPC_CUDA(cudaFuncSetAttribute(KERNEL<8, true, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                    KERNEL<8, true, IDX><<<grid, kLayerThreads, smem, st>>>(__VA_ARGS__); break;         \
        }                                                                                           \
    } while (0)

int ldpc_check_layer_fwd(const float* x, const int64_t* idx, int64_t B, int64_t E, int K, float* out,
                         int32_t* argmin_out, void* stream) {
    if (!x || !idx || !out) return fail(LDPC_ERR_INVALID, "check_layer_fwd: null argument");
    if (B < 0 || E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "check_layer_fwd: bad shape");
    if (B == 0) return LDPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_LAYER_DISPATCH(check_layer_fwd_kernel, IdxI64, x, IdxI64{(const long long*)idx}, (long long)B, (long long)E, K, out, argmin_out);
    LDPC_CHECK_LAUNCH("check_layer_fwd_kernel");
    return LDPC_OK;
}

int ldpc_check_layer_bwd(const float* x, const int64_t* idx, const int32_t* argmin, const float* grad_out, int64_t B,
                         int64_t E, int K, float* grad_x, void* stream) {
    if (!x || !idx || !argmin || !grad_out || !grad_x) return fail(LDPC_ERR_INVALID, "check_layer_bwd: null argument");
    if (B < 0 || E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "check_layer_bwd: bad shape");
    if (B == 0) return LDPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)B * E, st));
    check_layer_bwd_kernel<<<layer_grid(B * E, kLayerThreads), kLayerThreads, 0, st>>>(
        x, (const long long*)idx, argmin, grad_out, (long long)B, (long long)E, K, grad_x);
    LDPC_CHECK_LAUNCH("check_layer_bwd_kernel");
    return LDPC_OK;
}

int ldpc_variable_layer_fwd(const float* llr, const float* c2v, const int64_t* idx, int64_t B, int64_t E, int K,
                            float* out, void* stream) {
    if (!llr || !c2v || !idx || !out) return fail(LDPC_ERR_INVALID, "variable_layer_fwd: null argument");
    if (B < 0 || E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "variable_layer_fwd: bad shape");
    if (B == 0) return LDPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_LAYER_DISPATCH(variable_layer_fwd_kernel, IdxI64, llr, c2v, IdxI64{(const long long*)idx}, (long long)B, (long long)E, K, out);
    LDPC_CHECK_LAUNCH("variable_layer_fwd_kernel");
    return LDPC_OK;
}

int ldpc_variable_layer_bwd(const int64_t* idx, const float* grad_out, int64_t B, int64_t E, int K, float* grad_c2v,
                            void* stream) {
    if (!idx || !grad_out || !grad_c2v) return fail(LDPC_ERR_INVALID, "variable_layer_bwd: null argument");
    if (B < 0 || E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "variable_layer_bwd: bad shape");
    if (B == 0) return LDPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_CUDA(cudaMemsetAsync(grad_c2v, 0, sizeof(float) * (size_t)B * E, st));
    variable_layer_bwd_kernel<<<layer_grid(B * E, kLayerThreads), kLayerThreads, 0, st>>>(
        (const long long*)idx, grad_out, (long long)B, (long long)E, K, grad_c2v);
    LDPC_CHECK_LAUNCH("variable_layer_bwd_kernel");
    return LDPC_OK;
}

int ldpc_residual_layer_fwd(const float* llr, const float* c2v, const float* w_ch, const float* w_res,
                            const float* const* prev, int L, int64_t B, int64_t E, float* out, void* stream) {
    if (!llr || !c2v || !w_ch || !out || (L > 0 && (!prev || !w_res))) return fail(LDPC_ERR_INVALID, "residual_layer_fwd: null argument");
    if (L < 0 || L > kMaxResidual) return fail(LDPC_ERR_UNSUPPORTED, "residual_layer_fwd: depth %d outside 0..%d", L, kMaxResidual);
    if (B < 0 || E <= 0) return fail(LDPC_ERR_INVALID, "residual_layer_fwd: bad shape");
    if (B == 0) return LDPC_OK;
    ResidualPtrs rp{};
    for (int i = 0; i < L; ++i) {
        if (!prev[i]) return fail(LDPC_ERR_INVALID, "residual_layer_fwd: prev[%d] is null", i);
        rp.prev[i] = prev[i];
    }
    residual_layer_fwd_kernel<<<layer_grid(B * E, kLayerThreads), kLayerThreads, 0, (cudaStream_t)stream>>>(
        llr, c2v, w_ch, w_res, rp, L, (long long)B, (long long)E, out);
    LDPC_CHECK_LAUNCH("residual_layer_fwd_kernel");
    return LDPC_OK;
}

int ldpc_neural_variable_layer_fwd(const float* llr, const float* c2v, const int64_t* idx, const float* w_ch,
                                   const float* w_res, const float* const* prev, int L, int64_t B, int64_t E, int K,
                                   float* out, void* stream) {
    if (!llr || !c2v || !idx || !w_ch || !out || (L > 0 && (!prev || !w_res)))
        return fail(LDPC_ERR_INVALID, "neural_variable_layer_fwd: null argument");
    if (L < 0 || L > kMaxResidual) return fail(LDPC_ERR_UNSUPPORTED, "neural_variable_layer_fwd: depth %d outside 0..%d", L, kMaxResidual);
    if (B < 0 || E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "neural_variable_layer_fwd: bad shape");
    if (B == 0) return LDPC_OK;
    ResidualPtrs rp{};
    for (int i = 0; i < L; ++i) {
        if (!prev[i]) return fail(LDPC_ERR_INVALID, "neural_variable_layer_fwd: prev[%d] is null", i);
        rp.prev[i] = prev[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_LAYER_DISPATCH(neural_variable_fwd_kernel, IdxI64, llr, c2v, IdxI64{(const long long*)idx}, w_ch, w_res, rp, L,
                        (long long)B, (long long)E, K, out);
    LDPC_CHECK_LAUNCH("neural_variable_fwd_kernel");
    return LDPC_OK;
}

// Sorted-pack variants (csrc/neural.cuh): idx16 [K,E] uint16 compacted columns, cnt [E] uint8, perm [E] uint16 or NULL.
static int sorted_grid(int64_t B) {
    long long g = (B + kPackedRows - 1) / kPackedRows;
    return (int)(g > 2ll * kNumSMs ? 2ll * kNumSMs : g);
}

int ldpc_check_layer_fwd_sorted(const float* x, const uint16_t* idx16, int K, const uint8_t* cnt, const uint16_t* perm,
                                int64_t B, int64_t E, float* out, int32_t* nstar, void* stream) {
    if (!x || !idx16 || !cnt || !out) return fail(LDPC_ERR_INVALID, "check_layer_fwd_sorted: null argument");
    if (B < 0 || E <= 0 || K <= 0 || K > 255 || E >= 0xFFFF) return fail(LDPC_ERR_INVALID, "check_layer_fwd_sorted: bad shape");
    const size_t smem = (size_t)kPackedRows * E * sizeof(float);
    if (smem > (size_t)110 * 1024) return fail(LDPC_ERR_UNSUPPORTED, "check_layer_fwd_sorted: %lld edges exceed the staging tile", (long long)E);
    if (B == 0) return LDPC_OK;
    if (K == 9) {         // create_LLR_mapping on the 5G BG2 graphs
        LDPC_CUDA(cudaFuncSetAttribute(sorted_check_fwd_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sorted_check_fwd_kernel<9><<<sorted_grid(B), kPackedThreads, smem, (cudaStream_t)stream>>>(
            x, idx16, K, cnt, perm, (long long)B, (int)E, out, nstar);
    } else {
        LDPC_CUDA(cudaFuncSetAttribute(sorted_check_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sorted_check_fwd_kernel<0><<<sorted_grid(B), kPackedThreads, smem, (cudaStream_t)stream>>>(
            x, idx16, K, cnt, perm, (long long)B, (int)E, out, nstar);
    }
    LDPC_CHECK_LAUNCH("sorted_check_fwd_kernel");
    return LDPC_OK;
}

int ldpc_variable_layer_fwd_sorted(const float* llr, const float* c2v, const uint16_t* idx16, int K, const uint8_t* cnt,
                                   const uint16_t* perm, const float* w_ch, const float* w_res, const float* const* prev,
                                   int L, int64_t B, int64_t E, float* out, void* stream) {
    if (!llr || !c2v || !idx16 || !cnt || !out) return fail(LDPC_ERR_INVALID, "variable_layer_fwd_sorted: null argument");
    if (L < 0 || L > kMaxResidual) return fail(LDPC_ERR_UNSUPPORTED, "variable_layer_fwd_sorted: depth %d outside 0..%d", L, kMaxResidual);
    if (L > 0 && (!w_ch || !w_res || !prev)) return fail(LDPC_ERR_INVALID, "variable_layer_fwd_sorted: residual terms need w_ch, w_res and prev");
    if (B < 0 || E <= 0 || K <= 0 || K > 255 || E >= 0xFFFF) return fail(LDPC_ERR_INVALID, "variable_layer_fwd_sorted: bad shape");
    const size_t smem = (size_t)kPackedRows * E * sizeof(float);
    if (smem > (size_t)110 * 1024) return fail(LDPC_ERR_UNSUPPORTED, "variable_layer_fwd_sorted: %lld edges exceed the staging tile", (long long)E);
    if (B == 0) return LDPC_OK;
    ResidualPtrs rp{};
    for (int i = 0; i < L; ++i) {
        if (!prev[i]) return fail(LDPC_ERR_INVALID, "variable_layer_fwd_sorted: prev[%d] is null", i);
        rp.prev[i] = prev[i];
    }
    if (K == 22) {
        LDPC_CUDA(cudaFuncSetAttribute(sorted_variable_fwd_kernel<22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sorted_variable_fwd_kernel<22><<<sorted_grid(B), kPackedThreads, smem, (cudaStream_t)stream>>>(
            llr, c2v, idx16, K, cnt, perm, w_ch, w_res, rp, L, (long long)B, (int)E, out);
    } else {
        LDPC_CUDA(cudaFuncSetAttribute(sorted_variable_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sorted_variable_fwd_kernel<0><<<sorted_grid(B), kPackedThreads, smem, (cudaStream_t)stream>>>(
            llr, c2v, idx16, K, cnt, perm, w_ch, w_res, rp, L, (long long)B, (int)E, out);
    }
    LDPC_CHECK_LAUNCH("sorted_variable_fwd_kernel");
    return LDPC_OK;
}

int ldpc_check_layer_bwd_nstar(const float* x, const float* out, const int32_t* nstar, const float* grad_out, int64_t B,
                               int64_t E, float* grad_x, void* stream) {
    if (!x || !out || !nstar || !grad_out || !grad_x) return fail(LDPC_ERR_INVALID, "check_layer_bwd_nstar: null argument");
    if (B < 0 || E <= 0) return fail(LDPC_ERR_INVALID, "check_layer_bwd_nstar: bad shape");
    if (B == 0) return LDPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LDPC_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)B * E, st));
    check_layer_bwd_nstar_kernel<<<layer_grid(B * E, 256), 256, 0, st>>>(x, out, nstar, grad_out, (long long)B, (long long)E, grad_x);
    LDPC_CHECK_LAUNCH("check_layer_bwd_nstar_kernel");
    return LDPC_OK;
}

int ldpc_neural_pack_index(const int64_t* idx, int64_t E, int K, uint16_t* out, void* stream) {
    if (!idx || !out) return fail(LDPC_ERR_INVALID, "neural_pack_index: null argument");
    if (E <= 0 || K <= 0) return fail(LDPC_ERR_INVALID, "neural_pack_index: bad shape");
    if (E >= 0xFFFF) return fail(LDPC_ERR_UNSUPPORTED, "neural_pack_index: %lld edges do not fit 16-bit indices", (long long)E);
    neural_pack_index_kernel<<<layer_grid(E * K, 256), 256, 0, (cudaStream_t)stream>>>((const long long*)idx, (long long)E, K, out);
    LDPC_CHECK_LAUNCH("neural_pack_index_kernel");
    return LDPC_OK;
}

}  // extern "C"

struct NeuralArgs {
    const float* llr; const uint16_t* cidx; int Kc; const uint8_t* ccnt; const uint16_t* cperm;
    const uint16_t* vidx; int Kv; const uint8_t* vcnt; const uint16_t* vperm;
    const float* w_ch; const float* w_res; int L; int iters; int64_t B; int E; const float* gt; float* soft; float* max_loss;
};

template <int kRows, int KC, int KV>
static int launch_neural_k(const NeuralArgs& a, size_t smem, cudaStream_t st) {
    LDPC_CUDA(cudaFuncSetAttribute(neural_decode_kernel<kRows, KC, KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas_per_sm = smem * 2 + 8192 <= (size_t)227 * 1024 ? 2 : 1;
    long long grid = (a.B + kRows - 1) / kRows;
    if (grid > (long long)kNumSMs * ctas_per_sm) grid = (long long)kNumSMs * ctas_per_sm;
    neural_decode_kernel<kRows, KC, KV><<<(int)grid, kNeuralThreads, smem, st>>>(
        a.llr, a.cidx, a.Kc, a.ccnt, a.cperm, a.vidx, a.Kv, a.vcnt, a.vperm, a.w_ch, a.w_res, a.L, a.iters, (long long)a.B, a.E, a.gt,
        a.soft, a.max_loss);
    LDPC_CHECK_LAUNCH("neural_decode_kernel");
    return LDPC_OK;
}

template <int kRows>
static int launch_neural(const NeuralArgs& a, size_t smem, cudaStream_t st) {
    if (a.Kc == 9 && a.Kv == 22)      // create_LLR_mapping on the 5G BG2 graphs (max degrees 10 / 23)
        return launch_neural_k<kRows, 9, 22>(a, smem, st);
    return launch_neural_k<kRows, 0, 0>(a, smem, st);
}

extern "C" {

int ldpc_neural_decode(const float* llr_e, const uint16_t* cidx, int Kc, const uint8_t* ccnt, const uint16_t* cperm,
                       const uint16_t* vidx, int Kv, const uint8_t* vcnt, const uint16_t* vperm, const float* w_ch, const float* w_res, int L, int iters, int64_t B, int64_t E,
                       const float* gt_e, float* soft, float* max_loss, void* stream) {
    if (!llr_e || !cidx || !ccnt || !vidx || !vcnt || !w_ch || !soft || (L > 0 && !w_res))
        return fail(LDPC_ERR_INVALID, "neural_decode: null argument");
    if (gt_e && !max_loss) return fail(LDPC_ERR_INVALID, "neural_decode: ground truth given without max_loss buffer");
    if (L < 0 || L > kNeuralMaxL) return fail(LDPC_ERR_UNSUPPORTED, "neural_decode: depth %d outside 0..%d", L, kNeuralMaxL);
    if (iters < 1 || B < 0 || E <= 0 || Kc <= 0 || Kv <= 0 || Kc > 255 || Kv > 255) return fail(LDPC_ERR_INVALID, "neural_decode: bad shape");
    if (E >= 0xFFFF) return fail(LDPC_ERR_UNSUPPORTED, "neural_decode: %lld edges do not fit 16-bit indices", (long long)E);
    if (B == 0) return LDPC_OK;
    const NeuralArgs a{llr_e, cidx, Kc, ccnt, cperm, vidx, Kv, vcnt, vperm, w_ch, w_res, L, iters, B, (int)E, gt_e, soft, max_loss};
    const size_t per_row = (size_t)((L > 0 ? L : 1) + 2) * E * sizeof(float);   // c2v + llr + ring
    const size_t cap = (size_t)220 * 1024;
    cudaStream_t st = (cudaStream_t)stream;
    if (4 * per_row <= cap && B >= 4 * 2 * kNumSMs) return launch_neural<4>(a, 4 * per_row, st);
    if (2 * per_row <= cap && B >= 2 * kNumSMs) return launch_neural<2>(a, 2 * per_row, st);
    if (per_row <= cap) return launch_neural<1>(a, per_row, st);
    return fail(LDPC_ERR_UNSUPPORTED, "neural_decode: %zu bytes of resident state per codeword exceed shared memory", per_row);
}

int ldpc_neural_decode_qc(const ldpc_code_t* code, const float* llr_e, const float* w_ch, const float* w_res, int L, int iters,
                          int64_t B, const float* gt_e, float* soft, float* max_loss, float* save_x, int32_t* argmax, void* stream) {
    if (!code || !llr_e || !w_ch || !soft || (L > 0 && !w_res)) return fail(LDPC_ERR_INVALID, "neural_decode_qc: null argument");
    if (gt_e && !max_loss) return fail(LDPC_ERR_INVALID, "neural_decode_qc: ground truth given without max_loss buffer");
    if (L < 0 || L > 2) return fail(LDPC_ERR_UNSUPPORTED, "neural_decode_qc: residual depth %d outside 0..2 (two ring slots in Tensor Memory)", L);
    if (iters < 1 || B < 0) return fail(LDPC_ERR_INVALID, "neural_decode_qc: bad shape");
    if (B == 0) return LDPC_OK;
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "neural_decode_qc: cannot select device %d", code->device);
    NeuralQcParams p{};
    p.llr = llr_e; p.w_ch = w_ch; p.w_res = w_res; p.L = L; p.iters = iters; p.B = B; p.gt = gt_e; p.soft = soft;
    p.max_loss = max_loss; p.save_x = save_x; p.argmax = argmax;
    if (argmax && !gt_e) return fail(LDPC_ERR_INVALID, "neural_decode_qc: argmax needs the ground truth");
    return launch_neural_qc(code, p, (cudaStream_t)stream);
}

int ldpc_neural_backward_qc(const ldpc_code_t* code, const float* save_x, const float* soft, const float* gt_e, const int32_t* argmax,
                            const float* g_ml, const float* w_res, int L, int iters, int64_t B, float* g_wch, float* g_wres,
                            void* stream) {
    if (!code || !save_x || !soft || !gt_e || !argmax || !g_ml || !g_wch || (L > 0 && (!w_res || !g_wres)))
        return fail(LDPC_ERR_INVALID, "neural_backward_qc: null argument");
    if (L < 0 || L > 2) return fail(LDPC_ERR_UNSUPPORTED, "neural_backward_qc: residual depth %d outside 0..2", L);
    if (iters < 1 || B < 0) return fail(LDPC_ERR_INVALID, "neural_backward_qc: bad shape");
    if (B == 0) return LDPC_OK;
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "neural_backward_qc: cannot select device %d", code->device);
    NeuralQcBwdParams p{};
    p.save_x = save_x; p.soft = soft; p.gt = gt_e; p.argmax = argmax; p.g_ml = g_ml; p.w_res = w_res; p.L = L; p.iters = iters;
    p.B = B; p.g_wch = g_wch; p.g_wres = g_wres;
    return launch_neural_qc_bwd(code, p, (cudaStream_t)stream);
}

int ldpc_neural_decode_qc_var(const ldpc_code_t* code, const float* llr_v, const float* w_ch, const float* w_res, int L, int iters,
                              int64_t B, const float* gt_v, float* soft_v, float* max_loss, float* save_x, int32_t* argmax,
                              float* star, void* stream) {
    if (!code || !llr_v || !w_ch || !soft_v || (L > 0 && !w_res)) return fail(LDPC_ERR_INVALID, "neural_decode_qc_var: null argument");
    if (gt_v && !max_loss) return fail(LDPC_ERR_INVALID, "neural_decode_qc_var: ground truth given without max_loss buffer");
    if ((argmax || star) && !gt_v) return fail(LDPC_ERR_INVALID, "neural_decode_qc_var: argmax / star need the ground truth");
    if (L < 0 || L > 2) return fail(LDPC_ERR_UNSUPPORTED, "neural_decode_qc_var: residual depth %d outside 0..2 (two ring slots in Tensor Memory)", L);
    if (iters < 1 || B < 0) return fail(LDPC_ERR_INVALID, "neural_decode_qc_var: bad shape");
    if (B == 0) return LDPC_OK;
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "neural_decode_qc_var: cannot select device %d", code->device);
    NeuralQcParams p{};
    p.llr = llr_v; p.w_ch = w_ch; p.w_res = w_res; p.L = L; p.iters = iters; p.B = B; p.gt = gt_v; p.soft = soft_v;
    p.max_loss = max_loss; p.save_x = save_x; p.argmax = argmax; p.per_var = 1; p.star = star;
    return launch_neural_qc(code, p, (cudaStream_t)stream);
}

int ldpc_neural_backward_qc_var(const ldpc_code_t* code, const float* save_x, const float* star, const int32_t* argmax,
                                const float* g_ml, const float* w_res, int L, int iters, int64_t B, float* g_wch, float* g_wres,
                                void* stream) {
    if (!code || !save_x || !star || !argmax || !g_ml || !g_wch || (L > 0 && (!w_res || !g_wres)))
        return fail(LDPC_ERR_INVALID, "neural_backward_qc_var: null argument");
    if (L < 0 || L > 2) return fail(LDPC_ERR_UNSUPPORTED, "neural_backward_qc_var: residual depth %d outside 0..2", L);
    if (iters < 1 || B < 0) return fail(LDPC_ERR_INVALID, "neural_backward_qc_var: bad shape");
    if (B == 0) return LDPC_OK;
    DeviceGuard g(code->device);
    if (!g.ok) return fail(LDPC_ERR_CUDA, "neural_backward_qc_var: cannot select device %d", code->device);
    NeuralQcBwdParams p{};
    p.save_x = save_x; p.star = star; p.argmax = argmax; p.g_ml = g_ml; p.w_res = w_res; p.L = L; p.iters = iters;
    p.B = B; p.g_wch = g_wch; p.g_wres = g_wres;
    return launch_neural_qc_bwd(code, p, (cudaStream_t)stream);
}

int ldpc_output_layer_fwd(const float* final_llr, const float* llr, const float* gt, int64_t B, int64_t E, float* soft,
                          float* max_loss, int32_t* argmax, void* stream) {
    if (!final_llr || !llr || !soft) return fail(LDPC_ERR_INVALID, "output_layer_fwd: null argument");
    if (gt && !max_loss) return fail(LDPC_ERR_INVALID, "output_layer_fwd: ground truth given without max_loss buffer");
    if (B < 0 || E <= 0) return fail(LDPC_ERR_INVALID, "output_layer_fwd: bad shape");
    if (B == 0) return LDPC_OK;
    output_layer_fwd_kernel<<<layer_grid(B, kLayerThreads / 32), kLayerThreads, 0, (cudaStream_t)stream>>>(
        final_llr, llr, gt, (long long)B, (long long)E, soft, max_loss, argmax);
    LDPC_CHECK_LAUNCH("output_layer_fwd_kernel");
    return LDPC_OK;
}

// ---- message-centred GNN -------------------------------------------------------------------
static size_t gnn_bytes_per_codeword(const ldpc_gnn_t* g) {
    return (size_t)(2 * (size_t)g->E + g->N + g->M) * kH * sizeof(float);
}

int ldpc_gnn_create(const ldpc_code_t* code, int num_layers, int hidden, int num_types, const int32_t* edge_type,
                    ldpc_gnn_t** out) {
    if (!out) return fail(LDPC_ERR_INVALID, "gnn_create: null out pointer");
    *out = nullptr;
    if (!code) return fail(LDPC_ERR_INVALID, "gnn_create: null code handle");
    if (hidden != kH) return fail(LDPC_ERR_UNSUPPORTED, "gnn_create: hidden_dim %d (kernels are built for %d)", hidden, kH);
    if (num_layers < 1 || num_layers > 64) return fail(LDPC_ERR_INVALID, "gnn_create: num_layers %d", num_layers);
    if (num_types < 1) return fail(LDPC_ERR_INVALID, "gnn_create: num_types %d", num_types);
    const int Z = code->Z, rows = code->rows, cols = code->cols;
    ldpc_gnn* g = new (std::nothrow) ldpc_gnn();
    if (!g) return fail(LDPC_ERR_NOMEM, "gnn_create: out of host memory");
    g->code = code; g->device = code->device; g->layers = num_layers; g->hidden = hidden; g->types = num_types;
    g->E = code->E * Z; g->N = code->N; g->M = code->M;
    GnnLayout lay{num_types};
    g->params = (size_t)lay.total(num_layers);
    // message list in the reference's order (message_gnn_decoder.py:397-406): check-major, ascending variable
    std::vector<int> ev(g->E), ec(g->E), et(g->E), cptr(g->M + 1), vptr(g->N + 1, 0), vedge(g->E);
    const uint32_t* t = code->h_tab.data();
    const int off_rowptr = t[7], off_redge = t[9];
    int e = 0;
    for (int i = 0; i < rows; ++i)
        for (int r = 0; r < Z; ++r) {
            cptr[i * Z + r] = e;
            for (int k = (int)t[off_rowptr + i]; k < (int)t[off_rowptr + i + 1]; ++k) {
                const int j = t[off_redge + k] & 0xffff, s = (t[off_redge + k] >> 16) & 0xff;
                ev[e] = j * Z + (r + s) % Z;
                ec[e] = i * Z + r;
                int ty = edge_type ? edge_type[k] : 0;
                ty = ty < 0 ? 0 : (ty >= num_types ? num_types - 1 : ty);      // clamp, as :81
                et[e] = ty;
                ++e;
            }
        }
    cptr[g->M] = e;
    (void)cols;
    for (int x = 0; x < g->E; ++x) vptr[ev[x] + 1]++;
    for (int v = 0; v < g->N; ++v) vptr[v + 1] += vptr[v];
    std::vector<int> fill(g->N, 0);
    for (int x = 0; x < g->E; ++x) vedge[vptr[ev[x]] + fill[ev[x]]++] = x;
    // packed lists of the pipelined node kernel: message index and message type in one entry
    const bool node_pipe = g->E < (1 << kNpListShift) && num_types <= kNpMaxTypes && g->N % 4 == 0 && g->M % 4 == 0;
    std::vector<int> vlist2(node_pipe ? g->E : 0), clist2(node_pipe ? g->E : 0);
    for (int x = 0; node_pipe && x < g->E; ++x) {
        vlist2[x] = vedge[x] | (et[vedge[x]] << kNpListShift);
        clist2[x] = x | (et[x] << kNpListShift);
    }
    DeviceGuard dg(g->device);
    if (!dg.ok) { delete g; return fail(LDPC_ERR_CUDA, "gnn_create: cannot select device"); }
    auto up = [&](int** dst, const std::vector<int>& src) -> bool {
        if (cudaMalloc(dst, sizeof(int) * src.size()) != cudaSuccess) return false;
        return cudaMemcpy(*dst, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    bool ok = up(&g->d_edge_var, ev) && up(&g->d_edge_chk, ec) && up(&g->d_edge_type, et) && up(&g->d_var_ptr, vptr) &&
              up(&g->d_var_edge, vedge) && up(&g->d_chk_ptr, cptr) && (!node_pipe || (up(&g->d_var_list2, vlist2) && up(&g->d_chk_list2, clist2))) &&
              cudaMalloc(&g->d_packed, sizeof(float) * (size_t)num_layers * kPackedPerLayer) == cudaSuccess &&
              cudaMalloc(&g->d_emb, sizeof(float) * (size_t)num_layers * num_types * kH) == cudaSuccess &&
              cudaMalloc(&g->d_tc, sizeof(float) * (size_t)num_layers * kTcPerLayer) == cudaSuccess &&
              cudaMalloc(&g->d_tc16, sizeof(__half) * (size_t)num_layers * kTc16PerLayer) == cudaSuccess &&
              cudaMalloc(&g->d_status, sizeof(int)) == cudaSuccess && cudaMemset(g->d_status, 0, sizeof(int)) == cudaSuccess;
    if (!ok) { ldpc_gnn_destroy(g); return fail(LDPC_ERR_CUDA, "gnn_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); }
    *out = g;
    return LDPC_OK;
}

int ldpc_gnn_destroy(ldpc_gnn_t* g) {
    if (!g) return LDPC_OK;
    {
        DeviceGuard dg(g->device);
        cudaFree(g->d_edge_var); cudaFree(g->d_edge_chk); cudaFree(g->d_edge_type); cudaFree(g->d_var_ptr);
        cudaFree(g->d_var_edge); cudaFree(g->d_chk_ptr); cudaFree(g->d_var_list2); cudaFree(g->d_chk_list2); cudaFree(g->d_packed); cudaFree(g->d_emb); cudaFree(g->d_tc); cudaFree(g->d_tc16); cudaFree(g->d_status);
    }
    delete g;
    return LDPC_OK;
}

size_t ldpc_gnn_param_count(const ldpc_gnn_t* g) { return g ? g->params : 0; }

// Training workspace layout (floats), B codewords, L layers:
//   X[L+1][B][E][h]  PV[L][B][N][h]  PC[L][B][M][h]  SOFT[B][N]  DSOFT[B][N]
//   G[B][E][h] DC[B][E][h] HR[B][E][2h] DH[B][E][2h]
//   DPV,MV,DMV [B][N][h]   DPC,MC,DMC [B][M][h]   PG[L][kPackedPerLayer + types*h]
struct GnnTrainWs {
    size_t X, PV, PC, SOFT, DSOFT, G, DC, HR, DH, DPV, MV, DMV, DPC, MC, DMC, PG, MVS, MCS, total;   // MVS/MCS: node means of every layer
    size_t xs, pvs, pcs, pg_layer;
};
static GnnTrainWs gnn_train_layout(const ldpc_gnn_t* g, int64_t B) {
    GnnTrainWs w{};
    const size_t L = g->layers, E = g->E, N = g->N, M = g->M, b = (size_t)B;
    w.xs = b * E * kH; w.pvs = b * N * kH; w.pcs = b * M * kH; w.pg_layer = (size_t)kPackedPerLayer + (size_t)g->types * kH;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
    w.X = take((L + 1) * w.xs); w.PV = take(L * w.pvs); w.PC = take(L * w.pcs);
    w.SOFT = take(b * N); w.DSOFT = take(b * N);
    w.G = take(w.xs); w.DC = take(w.xs); w.HR = take(2 * w.xs); w.DH = take(2 * w.xs);
    w.DPV = take(w.pvs); w.MV = take(w.pvs); w.DMV = take(w.pvs);
    w.DPC = take(w.pcs); w.MC = take(w.pcs); w.DMC = take(w.pcs);
    w.PG = take(L * w.pg_layer);
    w.MVS = take(L * w.pvs); w.MCS = take(L * w.pcs);
    w.total = o;
    return w;
}

size_t ldpc_gnn_workspace_bytes(const ldpc_gnn_t* g, int64_t B, int training) {
    if (!g || B <= 0) return 0;
    if (training) return gnn_train_layout(g, B).total * sizeof(float);
    const int64_t chunk = B < 2048 ? B : 2048;
    return gnn_bytes_per_codeword(g) * (size_t)chunk;
}

int ldpc_gnn_forward(const ldpc_gnn_t* g, const float* params, const float* llr, int64_t B, float* soft_out,
                     float* prob_out, void* workspace, size_t ws_bytes, int training, void* stream) {
    if (!g || !params) return fail(LDPC_ERR_INVALID, "gnn_forward: null handle or parameters");
    if (B < 0) return fail(LDPC_ERR_INVALID, "gnn_forward: negative batch");
    if (B == 0) return LDPC_OK;
    if (!llr) return fail(LDPC_ERR_INVALID, "gnn_forward: null llr");
    const size_t per_cw = gnn_bytes_per_codeword(g);
    const size_t need = training ? gnn_train_layout(g, B).total * sizeof(float) : per_cw;
    if (!workspace || ws_bytes < need)
        return fail(LDPC_ERR_INVALID, "gnn_forward: workspace of %zu bytes, need at least %zu", ws_bytes, need);
    DeviceGuard dg(g->device);
    if (!dg.ok) return fail(LDPC_ERR_CUDA, "gnn_forward: cannot select device %d", g->device);
    cudaStream_t st = (cudaStream_t)stream;
    GnnLayout lay{g->types};
    gnn_pack_kernel<<<dim3(32, g->layers), 256, 0, st>>>(params, lay, g->layers, g->d_packed, g->d_emb);
    LDPC_CHECK_LAUNCH("gnn_pack_kernel");
    const size_t edge_smem = sizeof(float) * kEdgeSmemFloats;
    LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem));
    LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem));
    // tensor-core path (tcgen05, 3xTF32) unless LDPC_GNN_FFMA=1 asks for the fp32 FFMA kernels
    static const bool use_tc = !(getenv("LDPC_GNN_FFMA") && getenv("LDPC_GNN_FFMA")[0] == '1');
    // LDPC_GNN_EDGE=serial selects the single-buffered edge kernel (gnn_tc.cuh) instead of the pipelined one (diagnostics)
    static const bool use_pipe = !(getenv("LDPC_GNN_EDGE") && getenv("LDPC_GNN_EDGE")[0] == 's');
    if (use_tc) {
        gnn_pack_tc_kernel<<<dim3(16, g->layers), 256, 0, st>>>(g->d_packed, g->d_tc, (__half*)g->d_tc16);
        LDPC_CHECK_LAUNCH("gnn_pack_tc_kernel");
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdgeTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdgeTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_node_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNodeTcSmem));
        if (g->d_var_list2)
            LDPC_CUDA(cudaFuncSetAttribute(gnn_node_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node_pipe_smem(g->types)));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_pipe_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
    }
    auto tc_grid = [](long long rows, int per_sm) {
        const long long tiles = (rows + 127) / 128, cap = (long long)kNumSMs * per_sm;
        return (int)(tiles < cap ? tiles : cap);
    };
    const int E = g->E, N = g->N, M = g->M;
    const GnnTrainWs tw = training ? gnn_train_layout(g, B) : GnnTrainWs{};
    const int64_t chunk_max = training ? B : (int64_t)(ws_bytes / per_cw);
    for (int64_t b0 = 0; b0 < B; b0 += chunk_max) {
        const int64_t bc = (B - b0) < chunk_max ? (B - b0) : chunk_max;
        float* base = (float*)workspace;
        float* xa = training ? base + tw.X : base;
        float* xb = training ? xa + tw.xs : xa + (size_t)bc * E * kH;
        float* pv = training ? base + tw.PV : xb + (size_t)bc * E * kH;
        float* pc = training ? base + tw.PC : pv + (size_t)bc * N * kH;
        const float* llr_c = llr + (size_t)b0 * N;
        bool fused_readout = false;
        gnn_embed_kernel<<<gnn_grid(bc * E * (kH / 4), 256), 256, 0, st>>>(params, lay, llr_c, g->d_edge_var, bc, E, N, xa);
        LDPC_CHECK_LAUNCH("gnn_embed_kernel");
        for (int l = 0; l < g->layers; ++l) {
            const float* pk = g->d_packed + (size_t)l * kPackedPerLayer;
            const float* em = g->d_emb + (size_t)l * g->types * kH;
            if (use_tc) {
                const float* tcw = g->d_tc + (size_t)l * kTcPerLayer;
                float* mvs = training ? base + tw.MVS + (size_t)l * tw.pvs : nullptr;
                float* mcs = training ? base + tw.MCS + (size_t)l * tw.pcs : nullptr;
                // node terms: the warp-specialised gather pipeline (gnn_node_pipe.cuh); LDPC_GNN_NODE=serial selects the
                // phase-by-phase kernel it replaced (bit-identical, kept for codes whose tables do not fit the packed lists)
                static const bool node_serial = getenv("LDPC_GNN_NODE") && getenv("LDPC_GNN_NODE")[0] == 's';
                if (g->d_var_list2 && !node_serial) {
                    const size_t nsm = node_pipe_smem(g->types);
                    gnn_node_pipe_kernel<<<tc_grid(bc * N, 1), kNpThreads, nsm, st>>>(
                        xa, em, pk, tcw, 0, g->d_var_ptr, g->d_var_list2, g->types, bc, E, N, pv, mvs, g->d_status);
                    LDPC_CHECK_LAUNCH("gnn_node_pipe_kernel(var)");
                    gnn_node_pipe_kernel<<<tc_grid(bc * M, 1), kNpThreads, nsm, st>>>(
                        xa, em, pk, tcw, 1, g->d_chk_ptr, g->d_chk_list2, g->types, bc, E, M, pc, mcs, g->d_status);
                    LDPC_CHECK_LAUNCH("gnn_node_pipe_kernel(chk)");
                } else {
                    gnn_node_tc_kernel<<<tc_grid(bc * N, 2), kNodeThreads, kNodeTcSmem, st>>>(
                        xa, em, pk, tcw, 0, g->d_var_ptr, g->d_var_edge, g->d_edge_type, bc, E, N, pv, mvs, g->d_status);
                    LDPC_CHECK_LAUNCH("gnn_node_tc_kernel(var)");
                    gnn_node_tc_kernel<<<tc_grid(bc * M, 2), kNodeThreads, kNodeTcSmem, st>>>(
                        xa, em, pk, tcw, 1, g->d_chk_ptr, nullptr, g->d_edge_type, bc, E, M, pc, mcs, g->d_status);
                    LDPC_CHECK_LAUNCH("gnn_node_tc_kernel(chk)");
                }
                // inference: the last layer applies the readout projection itself and writes one float per message into xb
                fused_readout = use_pipe && !training && l == g->layers - 1;
                const float* wo = fused_readout ? params + lay.out_w(l) : nullptr;
                float* dec = fused_readout ? xb : nullptr;
                // edge MLPs: 3xTF32 operands by default; LDPC_GNN_MMA=f16 selects the fp16 two-way split (gnn_tc_pipe.cuh: twice
                // the tensor-pipe rate, half the operand footprint, 9.6e-6 instead of 1.8e-5 against the fp64 oracle -- and the
                // SAME 28.0 ms per forward at B = 2048: the tile time is set by the row warps' instruction stream, not by the
                // MMAs, so the range-safe form stays the default)
                static const bool use_f16 = getenv("LDPC_GNN_MMA") && getenv("LDPC_GNN_MMA")[0] == 'f';
                const float* tc16 = reinterpret_cast<const float*>((const __half*)g->d_tc16 + (size_t)l * kTc16PerLayer);
                if (use_pipe && use_f16 && l == 0)
                    gnn_edge_pipe_kernel<false, true><<<tc_grid(bc * E, 1), kPipeThreads, kPipeSmem, st>>>(
                        xa, em, pk, tc16, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, wo, dec, g->d_status);
                else if (use_pipe && use_f16)
                    gnn_edge_pipe_kernel<true, true><<<tc_grid(bc * E, 1), kPipeThreads, kPipeSmem, st>>>(
                        xa, em, pk, tc16, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, wo, dec, g->d_status);
                else if (use_pipe && l == 0)
                    gnn_edge_pipe_kernel<false, false><<<tc_grid(bc * E, 1), kPipeThreads, kPipeSmem, st>>>(
                        xa, em, pk, tcw, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, wo, dec, g->d_status);
                else if (use_pipe)
                    gnn_edge_pipe_kernel<true, false><<<tc_grid(bc * E, 1), kPipeThreads, kPipeSmem, st>>>(
                        xa, em, pk, tcw, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, wo, dec, g->d_status);
                else if (l == 0)
                    gnn_edge_tc_kernel<false><<<tc_grid(bc * E, 1), kEdgeThreads, kEdgeTcSmem, st>>>(
                        xa, em, pk, tcw, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, g->d_status);
                else
                    gnn_edge_tc_kernel<true><<<tc_grid(bc * E, 1), kEdgeThreads, kEdgeTcSmem, st>>>(
                        xa, em, pk, tcw, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb, g->d_status);
                LDPC_CHECK_LAUNCH("gnn_edge_tc_kernel");
            } else {
            gnn_node_kernel<<<gnn_grid(bc * N, kGnnThreads), kGnnThreads, 0, st>>>(
                xa, em, pk, 0, g->d_var_ptr, g->d_var_edge, g->d_edge_type, bc, E, N, pv);
            LDPC_CHECK_LAUNCH("gnn_node_kernel(var)");
            gnn_node_kernel<<<gnn_grid(bc * M, kGnnThreads), kGnnThreads, 0, st>>>(
                xa, em, pk, 1, g->d_chk_ptr, nullptr, g->d_edge_type, bc, E, M, pc);
            LDPC_CHECK_LAUNCH("gnn_node_kernel(chk)");
            if (l == 0)
                gnn_edge_kernel<false><<<gnn_grid(bc * E, kGnnThreads), kGnnThreads, edge_smem, st>>>(
                    xa, em, pk, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb);
            else
                gnn_edge_kernel<true><<<gnn_grid(bc * E, kGnnThreads), kGnnThreads, edge_smem, st>>>(
                    xa, em, pk, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, bc, E, N, M, xb);
            LDPC_CHECK_LAUNCH("gnn_edge_kernel");
            }
            if (training) { xa = xb; xb = xa + tw.xs; pv += tw.pvs; pc += tw.pcs; }
            else { float* tmp = xa; xa = xb; xb = tmp; }
        }
        float* soft_dst = training ? base + tw.SOFT : (soft_out ? soft_out + (size_t)b0 * N : nullptr);
        if (fused_readout) {
            gnn_readout_sum_kernel<<<gnn_grid(bc * N, 256), 256, 0, st>>>(
                xa, params, lay, g->layers - 1, llr_c, g->d_var_ptr, g->d_var_edge, bc, E, N, soft_dst,
                prob_out ? prob_out + (size_t)b0 * N : nullptr);
            LDPC_CHECK_LAUNCH("gnn_readout_sum_kernel");
        } else {
            gnn_readout_kernel<<<gnn_grid(bc * N, 256), 256, 0, st>>>(
                xa, params, lay, g->layers - 1, llr_c, g->d_var_ptr, g->d_var_edge, bc, E, N, soft_dst,
                prob_out ? prob_out + (size_t)b0 * N : nullptr);
            LDPC_CHECK_LAUNCH("gnn_readout_kernel");
        }
        if (training && soft_out)
            LDPC_CUDA(cudaMemcpyAsync(soft_out, soft_dst, sizeof(float) * (size_t)B * N, cudaMemcpyDeviceToDevice, st));
    }
    return LDPC_OK;
}

int ldpc_gnn_backward(const ldpc_gnn_t* g, const float* params, const float* llr, const float* gt, int64_t B,
                      float* loss_out, float* grad_params, void* workspace, size_t ws_bytes, void* stream) {
    if (!g || !params || !llr || !gt || !loss_out || !grad_params)
        return fail(LDPC_ERR_INVALID, "gnn_backward: null argument");
    if (B <= 0) return fail(LDPC_ERR_INVALID, "gnn_backward: batch must be positive");
    const GnnTrainWs tw = gnn_train_layout(g, B);
    if (!workspace || ws_bytes < tw.total * sizeof(float))
        return fail(LDPC_ERR_INVALID, "gnn_backward: workspace of %zu bytes, need %zu (from a training=1 forward)", ws_bytes,
                    tw.total * sizeof(float));
    DeviceGuard dg(g->device);
    if (!dg.ok) return fail(LDPC_ERR_CUDA, "gnn_backward: cannot select device %d", g->device);
    cudaStream_t st = (cudaStream_t)stream;
    GnnLayout lay{g->types};
    const int E = g->E, N = g->N, M = g->M, L = g->layers;
    float* base = (float*)workspace;
    float *G = base + tw.G, *DC = base + tw.DC, *HR = base + tw.HR, *DH = base + tw.DH;
    float *DPV = base + tw.DPV, *MV = base + tw.MV, *DMV = base + tw.DMV, *DPC = base + tw.DPC, *MC = base + tw.MC, *DMC = base + tw.DMC;
    float* PG = base + tw.PG;
    // The packed weight images live in the handle and an inference forward with other weights (EMA copy, evaluation
    // pass) may have overwritten them since the training forward: re-pack from THIS call's parameters (two small
    // launches).  A handle is still single-stream: d_packed / d_tc are shared scratch.
    gnn_pack_kernel<<<dim3(32, L), 256, 0, st>>>(params, lay, L, g->d_packed, g->d_emb);
    LDPC_CHECK_LAUNCH("gnn_pack_kernel");
    gnn_pack_tc_kernel<<<dim3(16, L), 256, 0, st>>>(g->d_packed, g->d_tc, (__half*)g->d_tc16);
    LDPC_CHECK_LAUNCH("gnn_pack_tc_kernel");
    LDPC_CUDA(cudaMemsetAsync(PG, 0, sizeof(float) * (size_t)L * tw.pg_layer, st));
    LDPC_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    // loss and d(loss)/d(soft)
    gnn_loss_kernel<<<gnn_grid(B * N, 256), 256, 0, st>>>(base + tw.SOFT, gt, (long long)B * N, base + tw.DSOFT, loss_out);
    LDPC_CHECK_LAUNCH("gnn_loss_kernel");
    const float* xL = base + tw.X + (size_t)L * tw.xs;
    gnn_readout_bwd_kernel<<<gnn_grid(B * E, 256), 256, 0, st>>>(xL, params, lay, L - 1, base + tw.DSOFT, g->d_edge_var, B, E, N, G,
                                                                grad_params);
    LDPC_CHECK_LAUNCH("gnn_readout_bwd_kernel");
    const size_t ebwd_smem = sizeof(float) * (2 * kH * kH + kH * 2 * kH + kH * kGnnThreads);
    LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ebwd_smem));
    const size_t fin_smem = sizeof(float) * (size_t)g->types * kH;
    const int outer_grid = kNumSMs * 2;
    // data gradients on the tensor cores (gnn_bwd_tc.cuh) unless LDPC_GNN_FFMA=1; the weight images are those the
    // training forward packed from the same parameters
    static const bool bwd_tc = !(getenv("LDPC_GNN_FFMA") && getenv("LDPC_GNN_FFMA")[0] == '1');
    if (bwd_tc) {
        LDPC_CUDA(cudaFuncSetAttribute(gnn_edge_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdgeBwdTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_dcomb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDcombTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_outer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOuterTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_outer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOuterTcSmem));
        LDPC_CUDA(cudaFuncSetAttribute(gnn_node_dm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNodeTcSmem));
    }
    for (int l = L - 1; l >= 0; --l) {
        const float* x = base + tw.X + (size_t)l * tw.xs;
        const float* pv = base + tw.PV + (size_t)l * tw.pvs;
        const float* pc = base + tw.PC + (size_t)l * tw.pcs;
        const float* pk = g->d_packed + (size_t)l * kPackedPerLayer;
        const float* em = g->d_emb + (size_t)l * g->types * kH;
        float* pg = PG + (size_t)l * tw.pg_layer;
        LDPC_CUDA(cudaMemsetAsync(DPV, 0, sizeof(float) * tw.pvs, st));
        LDPC_CUDA(cudaMemsetAsync(DPC, 0, sizeof(float) * tw.pcs, st));
        if (bwd_tc) {
            const float* tcw = g->d_tc + (size_t)l * kTcPerLayer;
            const long long tiles = ((long long)B * E + 127) / 128;
            const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
            gnn_edge_bwd_tc_kernel<<<grid, kBwdThreads, kEdgeBwdTcSmem, st>>>(
                x, em, tcw, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, G, B, E, N, M, HR, DH, DPV, DPC, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_edge_bwd_tc_kernel");
            gnn_dcomb_tc_kernel<<<grid, kBwdThreads, kDcombTcSmem, st>>>(DH, tcw, (long long)B * E, DC, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_dcomb_tc_kernel");
        } else {
        gnn_edge_bwd_kernel<<<gnn_grid(B * E, kGnnThreads), kGnnThreads, ebwd_smem, st>>>(
            x, em, pk, g->d_edge_var, g->d_edge_chk, g->d_edge_type, pv, pc, G, B, E, N, M, HR, DH, DPV, DPC);
        LDPC_CHECK_LAUNCH("gnn_edge_bwd_kernel");
        gnn_dcomb_kernel<<<gnn_grid(B * E, kGnnThreads), kGnnThreads, 0, st>>>(DH, pk, (long long)B * E, DC);
        LDPC_CHECK_LAUNCH("gnn_dcomb_kernel");
        }
        if (bwd_tc) {
            const long long tiles = ((long long)B * E + 127) / 128;
            const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
            // dW2[n][k] += sum_r G[r][n] relu(h)[r][k]  (D[m = k][n]);   d(b2) += colsum(G)
            gnn_outer_tc_kernel<false><<<grid, kOuterThreads, kOuterTcSmem, st>>>(
                HR, G, (long long)B * E, nullptr, nullptr, E, pg + kPkW2, 1, 2 * kH, nullptr, pg + kPkB2, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_outer_tc_kernel(dW2)");
            // dW1A[n][k] += sum_r dH[r][n] comb[r][k]   (D[m = n][k]);   d(b1v|b1c) += colsum(dH)
            gnn_outer_tc_kernel<true><<<grid, kOuterThreads, kOuterTcSmem, st>>>(
                DH, x, (long long)B * E, em, g->d_edge_type, E, pg + kPkW1A, kH, 1, pg + kPkB1V, nullptr, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_outer_tc_kernel(dW1A)");
        } else {
        // dW2[n][k] += G^T . relu(h);   d(b2) += colsum(G)
        gnn_outer_kernel<kH, 2 * kH, 0><<<outer_grid, 256, 0, st>>>(G, HR, (long long)B * E, nullptr, nullptr, E, pg + kPkW2, pg + kPkB2);
        LDPC_CHECK_LAUNCH("gnn_outer_kernel(dW2)");
        // dW1A[n][k] += dH^T . comb;   d(b1v|b1c) += colsum(dH)
        gnn_outer_kernel<2 * kH, kH, 2><<<outer_grid, 256, 0, st>>>(DH, x, (long long)B * E, em, g->d_edge_type, E, pg + kPkW1A, pg + kPkB1V);
        LDPC_CHECK_LAUNCH("gnn_outer_kernel(dW1A)");
        }
        if (bwd_tc) {       // the means were saved by the training forward; dm = dP . W1B / deg on the tensor cores
            const float* tcw = g->d_tc + (size_t)l * kTcPerLayer;
            MV = base + tw.MVS + (size_t)l * tw.pvs;
            MC = base + tw.MCS + (size_t)l * tw.pcs;
            auto ngrid = [](long long rows) { const long long t = (rows + 127) / 128; return (int)(t < 2 * kNumSMs ? t : 2 * kNumSMs); };
            gnn_node_dm_tc_kernel<<<ngrid((long long)B * N), kNodeThreads, kNodeTcSmem, st>>>(DPV, tcw, 0, g->d_var_ptr, (long long)B * N, N, DMV, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_node_dm_tc_kernel(var)");
            gnn_node_dm_tc_kernel<<<ngrid((long long)B * M), kNodeThreads, kNodeTcSmem, st>>>(DPC, tcw, 1, g->d_chk_ptr, (long long)B * M, M, DMC, g->d_status);
            LDPC_CHECK_LAUNCH("gnn_node_dm_tc_kernel(chk)");
        } else {
        gnn_node_bwd_kernel<<<gnn_grid(B * N, kGnnThreads), kGnnThreads, 0, st>>>(
            x, em, pk, 0, g->d_var_ptr, g->d_var_edge, g->d_edge_type, DPV, B, E, N, MV, DMV);
        LDPC_CHECK_LAUNCH("gnn_node_bwd_kernel(var)");
        gnn_node_bwd_kernel<<<gnn_grid(B * M, kGnnThreads), kGnnThreads, 0, st>>>(
            x, em, pk, 1, g->d_chk_ptr, nullptr, g->d_edge_type, DPC, B, E, M, MC, DMC);
        LDPC_CHECK_LAUNCH("gnn_node_bwd_kernel(chk)");
        }
        gnn_outer_kernel<kH, kH, 0><<<outer_grid, 256, 0, st>>>(DPV, MV, (long long)B * N, nullptr, nullptr, E, pg + kPkW1BV, nullptr);
        LDPC_CHECK_LAUNCH("gnn_outer_kernel(dW1Bv)");
        gnn_outer_kernel<kH, kH, 0><<<outer_grid, 256, 0, st>>>(DPC, MC, (long long)B * M, nullptr, nullptr, E, pg + kPkW1BC, nullptr);
        LDPC_CHECK_LAUNCH("gnn_outer_kernel(dW1Bc)");
        if (l > 0)
            gnn_finish_bwd_kernel<true><<<gnn_grid(B * E * (kH / 4), 256), 256, fin_smem, st>>>(
                DC, DMV, DMC, G, g->d_edge_var, g->d_edge_chk, g->d_edge_type, B, E, N, M, g->types, DC, pg + kPackedPerLayer);
        else
            gnn_finish_bwd_kernel<false><<<gnn_grid(B * E * (kH / 4), 256), 256, fin_smem, st>>>(
                DC, DMV, DMC, G, g->d_edge_var, g->d_edge_chk, g->d_edge_type, B, E, N, M, g->types, DC, pg + kPackedPerLayer);
        LDPC_CHECK_LAUNCH("gnn_finish_bwd_kernel");
        float* tmp = G; G = DC; DC = tmp;      // dx of this layer is the output gradient of the layer below
    }
    gnn_embed_bwd_kernel<<<kNumSMs * 8, 256, 0, st>>>(G, llr, g->d_edge_var, B, E, N, lay, grad_params);
    LDPC_CHECK_LAUNCH("gnn_embed_bwd_kernel");
    gnn_unpack_grad_kernel<<<dim3(32, L), 256, 0, st>>>(PG, lay, (int)tw.pg_layer, grad_params);
    LDPC_CHECK_LAUNCH("gnn_unpack_grad_kernel");
    return LDPC_OK;
}

}  // extern "C"
