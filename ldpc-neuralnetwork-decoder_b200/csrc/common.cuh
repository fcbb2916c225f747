// common.cuh -- error plumbing, launch accounting and small device helpers shared by every
// kernel of the engine.  Nothing here is visible through the C ABI except via ldpc_b200.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/ldpc_b200.h"

namespace ldpc {

// ---- host: last-error string (per thread) and launch counter --------------------------
inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
inline std::atomic<uint64_t>& launch_counter() {
    static std::atomic<uint64_t> c{0};
    return c;
}
#define LDPC_COUNT_LAUNCH() (::ldpc::launch_counter().fetch_add(1, std::memory_order_relaxed))

#define LDPC_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::ldpc::fail(LDPC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                 \
    } while (0)

#define LDPC_CHECK_LAUNCH(name)                                                                       \
    do {                                                                                             \
        LDPC_COUNT_LAUNCH();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess)                                                                       \
            return ::ldpc::fail(LDPC_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

constexpr int kNumSMs = 148;               // B200
constexpr int kMaxSmemPerBlock = 232448;   // 227 KB opt-in limit on sm_100

// ---- device helpers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float u2f(uint32_t x) { return __uint_as_float(x); }

}  // namespace ldpc
