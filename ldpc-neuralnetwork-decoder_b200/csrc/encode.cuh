// encode.cuh -- systematic encoder for QC codes of the shape H = [A B 0; C D I] (utils/encoder.py builds the plan).
// One warp per codeword; the codeword's bits live in shared memory (one byte each) so that the circulant shifts are
// rotations by address, as in the exact decoder.  plan = [g, kb, words, core_rows[g], core_cols[g], ext_of_row[rows]],
// binv = B^-1 packed by rows ([g*Z][words] uint32, bit k of a row = column k).
#pragma once
#include "tables.cuh"

namespace ldpc {

template <bool kConst>
__global__ void __launch_bounds__(256) encode_kernel(const uint32_t* gtab, int slot, const uint8_t* __restrict__ info, long long B,
                                                      const int* __restrict__ plan, const uint32_t* __restrict__ binv,
                                                      uint8_t* __restrict__ out) {
    extern __shared__ uint8_t enc_smem[];
    const Tab<kConst> tab{gtab, slot};
    const int rows = tab[0], cols = tab[1], Z = tab[2];
    const int off_rowptr = tab[7], off_redge = tab[9];
    const int g = plan[0], kb = plan[1], words = plan[2];
    const int* core_rows = plan + 3;
    const int* core_cols = core_rows + g;
    const int* ext_of_row = core_cols + g;
    const int N = cols * Z, K = kb * Z, gz = g * Z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int per_warp = ((N + 3) & ~3) + 4 * words;                       // bits[N] | packed lambda words
    uint8_t* bits = enc_smem + (size_t)warp * per_warp;
    uint32_t* lam = reinterpret_cast<uint32_t*>(bits + ((N + 3) & ~3));
    // parity of row i at circulant position r over all its cells except column `skip`
    auto row_parity = [&](int i, int r, int skip) -> unsigned {
        unsigned p = 0;
        for (int e = (int)tab[off_rowptr + i]; e < (int)tab[off_rowptr + i + 1]; ++e) {
            const uint32_t w = tab[off_redge + e];
            const int j = (int)(w & 0xffffu), s = (int)((w >> 16) & 0xffu);
            if (j == skip) continue;
            int q = r + s;
            q -= q >= Z ? Z : 0;
            p ^= bits[j * Z + q];
        }
        return p & 1u;
    };
    for (long long cw = (long long)blockIdx.x * W + warp; cw < B; cw += (long long)gridDim.x * W) {
        for (int n = lane; n < K; n += 32) bits[n] = info[cw * K + n] ? 1 : 0;
        for (int n = K + lane; n < N; n += 32) bits[n] = 0;
        __syncwarp();
        // lambda = A s for the core rows, packed 32 per word by ballot
        for (int t0 = 0; t0 < words * 32; t0 += 32) {
            const int t = t0 + lane;
            unsigned l = 0;
            if (t < gz) l = row_parity(core_rows[t / Z], t % Z, -1);     // core parity bits are still zero here
            const unsigned word = __ballot_sync(0xffffffffu, l != 0);
            if (lane == 0) lam[t0 >> 5] = word;
        }
        __syncwarp();
        // core parity = B^-1 lambda
        for (int t = lane; t < gz; t += 32) {
            unsigned acc = 0;
            for (int w = 0; w < words; ++w) acc ^= binv[(size_t)t * words + w] & lam[w];
            const unsigned par = __popc(acc) & 1u;            // parity(popc a + popc b) = parity(popc(a ^ b))
            bits[core_cols[t / Z] * Z + t % Z] = (uint8_t)par;
        }
        __syncwarp();
        // extension parities: each extension row determines its own degree-1 column
        for (int i = 0; i < rows; ++i) {
            const int x = ext_of_row[i];
            if (x < 0) continue;
            for (int r = lane; r < Z; r += 32) bits[x * Z + r] = (uint8_t)row_parity(i, r, x);
        }
        __syncwarp();
        for (int n = lane; n < N; n += 32) out[cw * N + n] = bits[n];
        __syncwarp();
    }
}

inline int launch_encode(const ldpc_code* c, const uint8_t* info, long long B, const int* plan, int words, const uint32_t* binv,
                         uint8_t* out, cudaStream_t st) {
    const size_t per_warp = (size_t)((c->N + 3) & ~3) + 4 * (size_t)words;
    int W = 8;
    while (W > 1 && per_warp * W > (size_t)kMaxSmemPerBlock) W >>= 1;
    if (per_warp * W > (size_t)kMaxSmemPerBlock) return fail(LDPC_ERR_UNSUPPORTED, "encode: codeword does not fit shared memory");
    long long blocks = (B + W - 1) / W;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    const size_t smem = per_warp * W;
    if (c->slot >= 0) {
        LDPC_CUDA(cudaFuncSetAttribute(encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_kernel<true><<<(int)blocks, W * 32, smem, st>>>(c->d_tab, c->slot, info, B, plan, binv, out);
    } else {
        LDPC_CUDA(cudaFuncSetAttribute(encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_kernel<false><<<(int)blocks, W * 32, smem, st>>>(c->d_tab, 0, info, B, plan, binv, out);
    }
    LDPC_CHECK_LAUNCH("encode_kernel");
    return LDPC_OK;
}

}  // namespace ldpc
