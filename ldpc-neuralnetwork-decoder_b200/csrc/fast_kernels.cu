// fast_kernels.cu -- translation unit of the fully unrolled specialised kernels
// (decode_fast_kernel.cuh) and their launcher.  Separate from ldpc_b200.cu so the two compile
// in parallel (the unrolled kernels dominate build time).
#include "decode_fast.cuh"
#include "decode_fast_kernel.cuh"

namespace ldpc {

#ifndef LDPC_FAST_WARPS
#define LDPC_FAST_WARPS 8
#endif
constexpr int kFastWarps = LDPC_FAST_WARPS;

template <class BG, int kAlgo, bool kEarly = false>
inline int launch_fast_inst(DecodeParams p, cudaStream_t st) {
    constexpr int G = 32 / BG::kZ;
    constexpr size_t smem = fast_smem_bytes<BG>(kFastWarps);
    static_assert(smem <= (size_t)kMaxSmemPerBlock, "fast kernel shared memory");
    auto kern = decode_fast_kernel<BG, kFastWarps, kAlgo, kEarly>;
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.ngroups = (p.B + G - 1) / G;
    long long blocks = (p.ngroups + kFastWarps - 1) / kFastWarps;
    if (blocks > kNumSMs) blocks = kNumSMs;
    kern<<<(int)blocks, kFastWarps * 32, smem, st>>>(p);
    LDPC_COUNT_LAUNCH();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFuncAttributes a{};
        cudaFuncGetAttributes(&a, kern);
        return fail(LDPC_ERR_CUDA, "launch of decode_fast_kernel failed: %s (regs %d, max threads/block %d, static smem %zu, "
                    "max dynamic smem %d, requested %d threads + %zu B)", cudaGetErrorString(e), a.numRegs,
                    a.maxThreadsPerBlock, a.sharedSizeBytes, a.maxDynamicSharedSizeBytes, kFastWarps * 32, smem);
    }
    return LDPC_OK;
}

int launch_fast(const ldpc_code* c, int algo, const DecodeParams& p, cudaStream_t st) {
    if (p.stop_mode == LDPC_STOP_PER_CODEWORD) {
        if (c->fast_kind == 1 && algo == LDPC_ALGO_MINSUM) return launch_fast_inst<BG2Z32, LDPC_ALGO_MINSUM, true>(p, st);
        if (c->fast_kind == 1 && algo == LDPC_ALGO_BP) return launch_fast_inst<BG2Z32, LDPC_ALGO_BP, true>(p, st);
        if (c->fast_kind == 2 && algo == LDPC_ALGO_MINSUM) return launch_fast_inst<BG2Z4, LDPC_ALGO_MINSUM, true>(p, st);
        if (c->fast_kind == 3 && algo == LDPC_ALGO_MINSUM) return launch_fast_inst<BG2Z16, LDPC_ALGO_MINSUM, true>(p, st);
        if (c->fast_kind == 3 && algo == LDPC_ALGO_BP) return launch_fast_inst<BG2Z16, LDPC_ALGO_BP, true>(p, st);
        if (c->fast_kind == 4 && algo == LDPC_ALGO_MINSUM) return launch_fast_inst<BG2Z8, LDPC_ALGO_MINSUM, true>(p, st);
        return fail(LDPC_ERR_UNSUPPORTED, "fast path: no per-codeword early-exit kernel for this table / algorithm");
    }
    if (algo == LDPC_ALGO_MINSUM) {
        if (c->fast_kind == 1) return launch_fast_inst<BG2Z32, LDPC_ALGO_MINSUM>(p, st);
        if (c->fast_kind == 2) return launch_fast_inst<BG2Z4, LDPC_ALGO_MINSUM>(p, st);
        if (c->fast_kind == 3) return launch_fast_inst<BG2Z16, LDPC_ALGO_MINSUM>(p, st);
        if (c->fast_kind == 4) return launch_fast_inst<BG2Z8, LDPC_ALGO_MINSUM>(p, st);
    } else {
        if (c->fast_kind == 1) return launch_fast_inst<BG2Z32, LDPC_ALGO_BP>(p, st);
        if (c->fast_kind == 2) return launch_fast_inst<BG2Z4, LDPC_ALGO_BP>(p, st);
        if (c->fast_kind == 3) return launch_fast_inst<BG2Z16, LDPC_ALGO_BP>(p, st);
        if (c->fast_kind == 4) return launch_fast_inst<BG2Z8, LDPC_ALGO_BP>(p, st);
    }
    return fail(LDPC_ERR_UNSUPPORTED, "fast path: code is not one of the compiled tables");
}

}  // namespace ldpc
