// params.cuh -- launch parameters and small device helpers shared by the decode kernels.
#pragma once
#include "channel.cuh"
#include "common.cuh"

namespace ldpc {

struct DecodeParams {
    const uint32_t* gtab;
    int slot;
    const float* llr;
    long long B;
    int iters;
    float alpha;
    int stop_mode;
    float* soft_out;
    void* hard_out;
    int hard_dtype;
    uint8_t* syndrome_ok;
    int32_t* iters_out;
    unsigned long long* valid_mask;
    int mask_words;
    int floats_per_warp;
    long long ngroups;
    GenParams gen;                   // sim mode: LLRs come from the on-chip channel, not from `llr`
    unsigned long long* counters;    // sim mode: [bit errors, frame errors, frames, undetected]
};

// channel LLR of bit n of global frame `frame` (all-zero codeword), see channel.cuh
__device__ __forceinline__ float gen_llr(const GenParams& g, unsigned long long frame, int n) {
    float z[4];
    normal4(g.seed, frame, (uint32_t)(((n >> 7) << 5) | (n & 31)), z);
    const int comp = (n >> 5) & 3;
    const float zz = comp == 0 ? z[0] : comp == 1 ? z[1] : comp == 2 ? z[2] : z[3];
    return llr_from_noise(zz, 1.0f, g);
}

// per-lane partial counters -> one atomicAdd per warp and counter
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ void flush_counters(unsigned long long* counters, unsigned long long bits,
                                               unsigned long long fe, unsigned long long frames,
                                               unsigned long long und) {
    bits = warp_sum_u64(bits); fe = warp_sum_u64(fe); frames = warp_sum_u64(frames); und = warp_sum_u64(und);
    if ((threadIdx.x & 31) == 0 && frames) {
        atomicAdd(&counters[0], bits);
        atomicAdd(&counters[1], fe);
        atomicAdd(&counters[2], frames);
        atomicAdd(&counters[3], und);
    }
}

// compile-time loop: f(IC<B>{}), f(IC<B+1>{}), ... with the index usable in constant expressions
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

}  // namespace ldpc
