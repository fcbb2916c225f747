// rate_match.cuh -- 5G NR rate matching (3GPP TS 38.212 5.4.2: bit selection from the circular buffer + bit
// interleaving) and its receiver-side inverse, the punctured / rate-matched LLR layout the decoder consumes
// (the reference transmits only the all-zero, un-punctured codeword: trainer.py:86,
// comparative_evaluation.py:132).  The index tables are built once per configuration on the host
// (utils/rate_match.py); the kernels are pure gathers -- HBM-bound, one pass, coalesced on the output side.
#pragma once
#include "common.cuh"

namespace ldpc {

// out[b][t] = cw[b][sel[t]]            transmitted bit t of codeword b
__global__ void __launch_bounds__(256) rate_match_kernel(const uint8_t* __restrict__ cw, const int* __restrict__ sel,
                                                          long long B, int N, int E, uint8_t* __restrict__ out) {
    const long long total = B * (long long)E;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long b = x / E;
        const int t = (int)(x - b * E);
        out[x] = cw[b * N + sel[t]];
    }
}

// llr[b][n] = base[n] + sum_{k in [ptr[n], ptr[n+1])} rx[b][idx[k]]   (ascending k = ascending transmission index):
// soft-combines repetitions, leaves punctured / untransmitted positions at base[n] = 0 and filler positions at
// base[n] = the "known zero" LLR.
__global__ void __launch_bounds__(256) rate_recover_kernel(const float* __restrict__ rx, const int* __restrict__ ptr,
                                                            const int* __restrict__ idx, const float* __restrict__ base,
                                                            long long B, int N, int E, float* __restrict__ llr) {
    const long long total = B * (long long)N;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long b = x / N;
        const int n = (int)(x - b * N);
        float acc = base[n];
        const float* r = rx + b * E;
        for (int k = ptr[n]; k < ptr[n + 1]; ++k) acc = __fadd_rn(acc, r[idx[k]]);
        llr[x] = acc;
    }
}

}  // namespace ldpc
