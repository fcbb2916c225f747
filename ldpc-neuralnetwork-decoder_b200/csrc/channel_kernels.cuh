// channel_kernels.cuh -- stand-alone channel / error-count kernels (device helpers: channel.cuh).
// Included by ldpc_b200.cu only.
#pragma once
#include <cuda_fp16.h>
#include "channel.cuh"

namespace ldpc {

// one thread per Philox block: writes up to four LLRs
__global__ void __launch_bounds__(256) awgn_llr_kernel(const uint8_t* __restrict__ bits, long long B, long long N,
                                                        GenParams g, float* __restrict__ out) {
    const long long nblk_per_frame = ((N + 127) >> 7) << 5;
    const long long total = B * nblk_per_frame;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / nblk_per_frame;
        const uint32_t blk = (uint32_t)(t - b * nblk_per_frame);
        float z[4];
        normal4(g.seed, g.first_frame + (unsigned long long)b, blk, z);
#pragma unroll
        for (int comp = 0; comp < 4; ++comp) {
            const long long n = ((long long)(blk >> 5) << 7) + ((long long)comp << 5) + (blk & 31);
            if (n < N) {
                const float s = bits && bits[b * N + n] ? -g.amp : g.amp;
                out[b * N + n] = llr_from_noise(z[comp], s, g);
            }
        }
    }
}

// bit / frame error counters.  One warp per codeword.
__global__ void __launch_bounds__(256) count_errors_kernel(const void* __restrict__ hard, int hard_dtype,
                                                            const uint8_t* __restrict__ tx, long long B, long long N,
                                                            unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long bit_err = 0, frame_err = 0, frames = 0;
    const long long NW = (N + 31) >> 5;
    for (long long b = warp0; b < B; b += nwarps) {
        unsigned e = 0;
        if (hard_dtype == LDPC_HARD_PACKED) {
            const unsigned* h = (const unsigned*)hard + b * NW;
            for (long long w = lane; w < NW; w += 32) {
                unsigned word = h[w];
                if (tx) {
                    unsigned t = 0;
                    for (int k = 0; k < 32; ++k) {
                        const long long n = w * 32 + k;
                        if (n < N && tx[b * N + n]) t |= 1u << k;
                    }
                    word ^= t;
                }
                e += __popc(word);
            }
        } else {
            for (long long n = lane; n < N; n += 32) {
                const int hb = hard_dtype == LDPC_HARD_F32 ? (((const float*)hard)[b * N + n] != 0.0f)
                                                          : (((const uint8_t*)hard)[b * N + n] != 0);
                const int tb = tx ? (tx[b * N + n] != 0) : 0;
                e += (hb != tb);
            }
        }
        e = __reduce_add_sync(0xffffffffu, e);
        bit_err += e;
        frame_err += e != 0;
        frames += 1;
    }
    if (lane == 0 && frames) {
        atomicAdd(&counters[0], bit_err);
        atomicAdd(&counters[1], frame_err);
        atomicAdd(&counters[2], frames);
    }
}

// flag[0] |= 1 if any of x[0..n) is +-inf or NaN: one read-only pass (128-bit loads), one atomic per CTA at most
__global__ void __launch_bounds__(256) nonfinite_flag_kernel(const float* __restrict__ x, long long n, int* __restrict__ flag) {
    const long long n4 = n >> 2;
    unsigned bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
        bad |= ((v.x & 0x7f800000u) == 0x7f800000u) | ((v.y & 0x7f800000u) == 0x7f800000u) |
               ((v.z & 0x7f800000u) == 0x7f800000u) | ((v.w & 0x7f800000u) == 0x7f800000u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) bad |= (__float_as_uint(x[(n4 << 2) + threadIdx.x]) & 0x7f800000u) == 0x7f800000u;
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// quantised LLRs as transferred -> the fp32 values the decoder (and the reference) computes on
template <typename T>
__global__ void llr_dequant_kernel(const T* __restrict__ raw, float scale, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if constexpr (sizeof(T) == 1) v = (float)raw[i];
        else v = __half2float(raw[i]);
        out[i] = v * scale;
    }
}


}  // namespace ldpc
