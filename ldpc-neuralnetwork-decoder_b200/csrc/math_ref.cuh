// math_ref.cuh -- transcendental functions with the reference's rounding behaviour.
//
// The reference's sum-product decoder is torch.tanh / torch.atanh on CPU fp32
// (models/traditional_decoders.py:78,81), unclipped, so which inputs saturate to exactly
// 1.0f (-> atanh = inf) is part of its observable behaviour.  CUDA's tanhf/atanhf are
// ~1-2 ulp routines with a different saturation point, so the exact path evaluates both in
// double precision and rounds once to fp32 (correctly rounded except for double-rounding
// cases of probability ~2^-29).  Measured in the build container: torch's CPU tanh differs
// from this by 1 ulp on 0.39 % of inputs in [-10,10] and atanh on 0.08 % (DESIGN.md,
// "BP numerics").  The same definitions are restated in oracle/ldpc_oracle.c.
#pragma once
#include "common.cuh"

namespace ldpc {

__device__ __forceinline__ float tanh_ref(float x) { return (float)tanh((double)x); }
__device__ __forceinline__ float atanh_ref(float x) { return (float)atanh((double)x); }

// Out of line on purpose (decode_exact.cuh): inlined kMaxDc times each, the double-precision routines made the
// sum-product body so large that the kernel starved on instruction fetch (ncu bp_exact_r1: no_instruction 3.98 stalls
// per issue, issue active 16.9 %).
static __device__ __noinline__ float tanh_half_ref(float v) { return tanh_ref(v * 0.5f); }
static __device__ __noinline__ float two_atanh_ref(float p) { return 2.0f * atanh_ref(p); }

}  // namespace ldpc
