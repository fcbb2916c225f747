// channel.cuh -- BPSK-AWGN channel on the device (Philox4x32-10 + Box-Muller) and integer
// error counters.
//
// Replaces AWGNChannel.transmit (utils/channel.py:205-231) and compute_ber_fer (:156-190).
// Reference arithmetic kept: symbols = 1-2*bit; sigma = 1/sqrt(10^(snr_db/10)) in float64,
// used as an fp32 scalar; received = s + z*sigma; llr = (2*received)/sigma^2, all fp32.
// The reference draws z with torch.randn on the host; here z comes from a counter-based
// generator so that a frame's noise depends only on (seed, global frame index, bit index):
//   key     = (seed_lo, seed_hi)
//   counter = (frame_lo, frame_hi, blk, 0),   blk = ((n >> 7) << 5) | (n & 31)
//   the block's four outputs serve bits n with (n >> 5) & 3 = 0..3, i.e. one lane of a warp
//   gets four consecutive 32-wide columns from a single Philox call.
//   u = x*2^-32 + 2^-33 (cuRAND's open-interval map), z0 = sqrt(-2 ln u0) cos(2 pi u1),
//   z1 = sqrt(-2 ln u0) sin(2 pi u1), likewise (z2, z3) from (x2, x3).
// oracle/ldpc_oracle.c restates the same generator on the CPU.
#pragma once
#include "common.cuh"

namespace ldpc {

struct GenParams {
    int enabled;                 // 0: read LLRs from memory
    float sigma;                 // fp32(1/sqrt(snr_linear))
    float var;                   // fp32(sigma_d * sigma_d)
    float amp;                   // symbol amplitude: 1 (BPSK, AWGNChannel) or fp32(1/sqrt(2)) (QPSK component)
    unsigned long long seed;
    unsigned long long first_frame;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// four standard normals for (frame, blk).  Box-Muller on the special-function unit: ln u = lg2(u) * ln 2 (MUFU.LG2),
// sqrt via MUFU.RSQ, sin/cos of 2 pi u as MUFU.SIN / MUFU.COS of an argument reduced to [-pi, pi) (u - 0.5 is exact,
// the half turn is undone by the sign).  Absolute error of a normal ~1e-6 -- the generator was never bit-identical to
// the oracle's libm version (tests hold LLRs to 2e-5 of their scale) -- and a Philox block costs ~170 instructions
// instead of ~350 (the fused simulation kernel spends 13 blocks per lane and codeword on its channel).
__device__ __forceinline__ void normal4(unsigned long long seed, unsigned long long frame, uint32_t blk, float (&z)[4]) {
    uint32_t x[4];
    philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), blk, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    const float r0 = __fsqrt_rn(-1.3862943611198906f * __log2f(u01(x[0]))), r1 = __fsqrt_rn(-1.3862943611198906f * __log2f(u01(x[2])));
    float s0, c0, s1, c1;
    __sincosf(6.2831853071795865f * (u01(x[1]) - 0.5f), &s0, &c0);          // angle - pi: sin and cos both change sign
    __sincosf(6.2831853071795865f * (u01(x[3]) - 0.5f), &s1, &c1);
    z[0] = -(r0 * c0); z[1] = -(r0 * s0); z[2] = -(r1 * c1); z[3] = -(r1 * s1);
}

// reference arithmetic (utils/channel.py:224-229): received = s + z*sigma; llr = (2*received)/sigma^2, fp32
__device__ __forceinline__ float llr_from_noise(float z, float symbol, const GenParams& g) {
    const float received = __fadd_rn(symbol, __fmul_rn(z, g.sigma));
    return __fdiv_rn(__fmul_rn(2.0f, received), g.var);
}

}  // namespace ldpc
