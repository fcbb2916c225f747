// neural_qc.cu -- translation unit of the QC-structured LDPCNeuralDecoder kernels (neural_qc.cuh) and their launcher;
// separate from ldpc_b200.cu so that the fully unrolled bodies compile in parallel with the rest.
#include "neural_qc_kernel.cuh"
#include "tables.cuh"

namespace ldpc {

template <class BG>
static int launch_neural_qc_bg(const NeuralQcParams& p, cudaStream_t st) {
    auto kern = neural_qc_kernel<BG>;
    constexpr size_t smem = neural_qc_smem_bytes<BG>();
    static_assert(smem <= (size_t)kMaxSmemPerBlock, "neural_qc shared memory");
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int per_cta = kNqGroups * (32 / BG::kZ);                     // codewords resident per CTA
    long long blocks = (p.B + per_cta - 1) / per_cta;
    if (blocks > kNumSMs) blocks = kNumSMs;
    kern<<<(int)blocks, kNqThreads, smem, st>>>(p);
    LDPC_CHECK_LAUNCH("neural_qc_kernel");
    return LDPC_OK;
}

int launch_neural_qc(const ldpc_code_t* c, const NeuralQcParams& p, cudaStream_t st) {
    if (c->fast_kind == 1) return launch_neural_qc_bg<BG2Z32>(p, st);
    if (c->fast_kind == 3) return launch_neural_qc_bg<BG2Z16>(p, st);      // the reference's default --lifting_factor: two codewords per warp
    if (c->fast_kind == 4) return launch_neural_qc_bg<BG2Z8>(p, st);
    if (c->fast_kind == 2) return launch_neural_qc_bg<BG2Z4>(p, st);       // the shipped NR_2_0_4.txt: eight codewords per warp
    return fail(LDPC_ERR_UNSUPPORTED, "neural_decode_qc: the QC-structured kernel is compiled for the 5G BG2 tables at Z = 32, 16, 8, 4");
}

template <class BG>
static int launch_neural_qc_bwd_bg(const NeuralQcBwdParams& p, cudaStream_t st) {
    auto kern = neural_qc_bwd_kernel<BG>;
    constexpr size_t smem = neural_qc_bwd_smem_bytes<BG>();
    static_assert(smem <= (size_t)kMaxSmemPerBlock, "neural_qc_bwd shared memory");
    LDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int per_cta = kNqGroups * (32 / BG::kZ);
    long long blocks = (p.B + per_cta - 1) / per_cta;
    if (blocks > kNumSMs) blocks = kNumSMs;
    kern<<<(int)blocks, kNqThreads, smem, st>>>(p);
    LDPC_CHECK_LAUNCH("neural_qc_bwd_kernel");
    return LDPC_OK;
}

int launch_neural_qc_bwd(const ldpc_code_t* c, const NeuralQcBwdParams& p, cudaStream_t st) {
    if (c->fast_kind == 1) return launch_neural_qc_bwd_bg<BG2Z32>(p, st);
    if (c->fast_kind == 3) return launch_neural_qc_bwd_bg<BG2Z16>(p, st);
    if (c->fast_kind == 4) return launch_neural_qc_bwd_bg<BG2Z8>(p, st);
    if (c->fast_kind == 2) return launch_neural_qc_bwd_bg<BG2Z4>(p, st);
    return fail(LDPC_ERR_UNSUPPORTED, "neural_backward_qc: the QC-structured kernel is compiled for the 5G BG2 tables at Z = 32, 16, 8, 4");
}

}  // namespace ldpc
