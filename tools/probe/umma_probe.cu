// Standalone probe: D[128 x N] = A[128 x K] . B[N x K]^T with tcgen05.mma kind::tf32 (SS operands,
// no swizzle, K-major canonical layout), optional 3xTF32 split, checked against a CPU fp64 reference.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical no-swizzle K-major layout: element (r, k) of a [R x K] fp32 operand
//   byte offset = (r/8)*SBO + (k/4)*128 + (r%8)*16 + (k%4)*4,  SBO = (K/4)*128
__device__ __forceinline__ uint32_t canon_off(int r, int k, int K) { return (r >> 3) * (K / 4) * 128 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
    return d;                          // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int N, int K, bool kSplit3>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Ahi = smem;                       // 128 x K
    uint8_t* Alo = Ahi + 128 * K * 4;
    uint8_t* Bhi = Alo + 128 * K * 4;          // N x K
    uint8_t* Blo = Bhi + N * K * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage operands: thread r owns row r of A (and row r of B if r < N)
    for (int k = 0; k < K; ++k) {
        const float a = A[tid * K + k];
        const float ah = to_tf32(a), al = to_tf32(a - ah);
        *reinterpret_cast<float*>(Ahi + canon_off(tid, k, K)) = ah;
        *reinterpret_cast<float*>(Alo + canon_off(tid, k, K)) = al;
        if (tid < N) {
            const float b = B[tid * K + k];
            const float bh = to_tf32(b), bl = to_tf32(b - bh);
            *reinterpret_cast<float*>(Bhi + canon_off(tid, k, K)) = bh;
            *reinterpret_cast<float*>(Blo + canon_off(tid, k, K)) = bl;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        const uint32_t sbo = (K / 4) * 128, lbo = 128;
        int first = 1;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t koff = ks * 2 * 128;          // two 16-byte chunks per k-step
            const uint64_t ah = make_desc(smem_u32(Ahi) + koff, lbo, sbo), al = make_desc(smem_u32(Alo) + koff, lbo, sbo);
            const uint64_t bh = make_desc(smem_u32(Bhi) + koff, lbo, sbo), bl = make_desc(smem_u32(Blo) + koff, lbo, sbo);
            auto mma = [&](uint64_t da, uint64_t db) {
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            };
            if (kSplit3) { mma(al, bh); mma(ah, bl); }
            mma(ah, bh);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs (bounded spin: never hang the GPU)
    {
        uint32_t done = 0;
        long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > 20000000LL) { if (tid == 0) *status = 1; break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: thread i of warp w reads TMEM lane 32w+i, all N columns
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}


// TS variant: the A operand (hi and lo images) lives in Tensor Memory, written by tcgen05.st (thread i of warp w
// owns lane 32w+i = row; element k of the row in column base + k), B from shared memory as above.
template <int N, int K>
__global__ void __launch_bounds__(128) probe_ts(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Bhi = smem;                       // N x K
    uint8_t* Blo = Bhi + N * K * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    constexpr uint32_t kCols = (N + 2 * K) <= 256 ? 256 : 512;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < N)
        for (int k = 0; k < K; ++k) {
            const float b = B[tid * K + k];
            const float bh = to_tf32(b), bl = to_tf32(b - bh);
            *reinterpret_cast<float*>(Bhi + canon_off(tid, k, K)) = bh;
            *reinterpret_cast<float*>(Blo + canon_off(tid, k, K)) = bl;
        }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_hi = tmem + N, a_lo = tmem + N + K;      // columns [N, N+K) and [N+K, N+2K)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 4) {
        uint32_t h[4], l[4];
        for (int j = 0; j < 4; ++j) {
            const float a = A[tid * K + k0 + j];
            const float ah = to_tf32(a);
            h[j] = __float_as_uint(ah); l[j] = __float_as_uint(to_tf32(a - ah));
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a_hi + lane_base + k0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a_lo + lane_base + k0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sbo = (K / 4) * 128, lbo = 128;
        int first = 1;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t koff = ks * 2 * 128;
            const uint64_t bh = make_desc(smem_u32(Bhi) + koff, lbo, sbo), bl = make_desc(smem_u32(Blo) + koff, lbo, sbo);
            auto mma = [&](uint32_t ta, uint64_t db) {
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             :: "r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            };
            mma(a_lo + ks * 8, bh); mma(a_hi + ks * 8, bl); mma(a_hi + ks * 8, bh);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > 20000000LL) { if (tid == 0) *status = 1; break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + lane_base + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kCols) : "memory");
}

template <int N, int K, bool S, bool TS = false>
int run(const char* name) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4); cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)(128 + N) * K * 4 * 2;
    if (TS) {
        cudaFuncSetAttribute(probe_ts<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_ts<N, K><<<1, 128, smem>>>(dA, dB, dD, dS);
    } else {
        cudaFuncSetAttribute(probe<N, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<N, K, S><<<1, 128, smem>>>(dA, dB, dD, dS);
    }
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    printf("%s N=%d K=%d split3=%d: cuda=%s timeout=%d max_abs_err=%.3e (max |ref| %.2f) rel=%.2e\n", name, N, K, (int)S,
           cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
    return e != cudaSuccess;
}

// MN-major variant (the weight-gradient contraction): D[M x N] = sum_r A[r][m] * B[r][n], both operands stored row r
// first (r = the MMA's K index), i.e. MN-major.  Canonical no-swizzle MN-major layout (cute/atom/mma_traits_sm100.hpp:
// ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 4 tf32): 16-byte chunks of 4 consecutive MN elements, 8 consecutive K rows
// 16 B apart (one 128-byte core matrix), MN groups SBO apart, K groups LBO apart.  Here [chunk][r/8][r%8][4]:
// SBO = (R/8)*128, LBO = 128.  kSwap exchanges the two fields to test the other reading of the descriptor.
template <int M, int N, int R, bool kSwap, int kMode = 0>
__global__ void __launch_bounds__(128) probe_mn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Ahi = smem;                       // M x R
    uint8_t* Alo = Ahi + M * R * 4;
    uint8_t* Bhi = Alo + M * R * 4;            // N x R
    uint8_t* Blo = Bhi + N * R * 4;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // kMode 1: SWIZZLE_128B_BASE32B, "the only available smem layout for MN-major tf32 operands" (CUTLASS sm100_common.inl:92):
    // atom = 32 MN elements (128 B) x 4 K rows 128 B apart, 32-byte chunk index XOR (K row % 4)  [Swizzle<2,5,2> on bytes];
    // K groups of 4 rows 512 B apart, MN groups of 32 elements (R/4)*512 B apart
    auto off = [](int mn, int r) {
        if (kMode == 0) return (uint32_t)((mn >> 2) * (R / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (mn & 3) * 4);
        const int g = mn >> 5, c32 = (mn & 31) >> 3, w = mn & 7;
        return (uint32_t)(g * (R / 4) * 512 + (r >> 2) * 512 + (r & 3) * 128 + ((c32 ^ (r & 3)) << 5) + w * 4);
    };
    for (int r = tid; r < R; r += 128) {
        for (int m = 0; m < M; ++m) {
            const float a = A[r * M + m];
            const float ah = to_tf32(a);
            *reinterpret_cast<float*>(Ahi + off(m, r)) = ah;
            *reinterpret_cast<float*>(Alo + off(m, r)) = to_tf32(a - ah);
        }
        for (int n = 0; n < N; ++n) {
            const float b = B[r * N + n];
            const float bh = to_tf32(b);
            *reinterpret_cast<float*>(Bhi + off(n, r)) = bh;
            *reinterpret_cast<float*>(Blo + off(n, r)) = to_tf32(b - bh);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // a_major (bit 15) = b_major (bit 16) = 1: MN-major
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        const uint32_t mn_stride = kMode == 0 ? (R / 8) * 128 : (R / 4) * 512, k_stride = kMode == 0 ? 128 : 512;
        const uint32_t sbo = kSwap ? mn_stride : k_stride, lbo = kSwap ? k_stride : mn_stride;
        const uint64_t lt = kMode == 0 ? 0ull : (1ull << 61);        // layout_type 1 = SWIZZLE_128B_BASE32B
        int first = 1;
        for (int ks = 0; ks < R / 8; ++ks) {
            const uint32_t koff = ks * (kMode == 0 ? 128 : 1024);      // next 8 K rows
            const uint64_t ah = make_desc(smem_u32(Ahi) + koff, lbo, sbo) | lt, al = make_desc(smem_u32(Alo) + koff, lbo, sbo) | lt;
            const uint64_t bh = make_desc(smem_u32(Bhi) + koff, lbo, sbo) | lt, bl = make_desc(smem_u32(Blo) + koff, lbo, sbo) | lt;
            auto mma = [&](uint64_t da, uint64_t db) {
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            };
            mma(al, bh); mma(ah, bl); mma(ah, bh);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > 20000000LL) { if (tid == 0) *status = 1; break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}

template <int M, int N, int R, bool kSwap, int kMode = 0>
int run_mn(const char* name) {
    std::vector<float> A(R * M), B(R * N), D(M * N);
    srand(2);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4); cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)(M + N) * R * 4 * 2;
    cudaFuncSetAttribute(probe_mn<M, N, R, kSwap, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_mn<M, N, R, kSwap, kMode><<<1, 128, smem>>>(dA, dB, dD, dS);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) printf("  launch error: %s\n", cudaGetErrorString(le));
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < R; ++k) r += (double)A[k * M + m] * B[k * N + n];
            maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    {
        int nz = 0;
        for (auto v : D) nz += v != 0.0f;
        double r00 = 0, r01 = 0, r10 = 0;
        for (int k = 0; k < R; ++k) { r00 += (double)A[k * M] * B[k * N]; r01 += (double)A[k * M] * B[k * N + 1]; r10 += (double)A[k * M + 1] * B[k * N]; }
        printf("  nonzero %d of %d; D[0][0..1], D[1][0] = %.4f %.4f %.4f  ref %.4f %.4f %.4f\n", nz, M * N, D[0], D[1], D[N], r00, r01, r10);
    }
    printf("%s mode=%d M=%d N=%d R=%d swap=%d: cuda=%s timeout=%d max_abs_err=%.3e (max |ref| %.2f) rel=%.2e\n", name, kMode, M, N, R, (int)kSwap,
           cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
    return e != cudaSuccess;
}


// ---- fp16 two-way split (round 2 experiment): x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits, and
// hi.hi + hi.lo + lo.hi on kind::f16 runs at twice the TF32 MMA rate with operands half as wide (A: K/2 TMEM columns per image,
// B: half the shared-memory image).  TS form: A in Tensor Memory (lane = row, column c holds k = 2c in the low and k = 2c+1 in
// the high half), B K-major canonical no-swizzle for 16-bit elements: core matrix = 8 rows x 8 elements (16 B).
// `reps` > 1 repeats the MMA sequence (accumulating) so that clock64 around it measures the tensor-pipe rate.
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t canon_off16(int r, int k, int K) { return (r >> 3) * (K / 8) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

template <int N, int K, bool kF16>
__global__ void __launch_bounds__(128) probe_ts_rate(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status,
                                                      int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int EB = kF16 ? 2 : 4;                 // bytes per operand element
    uint8_t* Bhi = smem;                             // N x K
    uint8_t* Blo = Bhi + N * K * EB;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    constexpr int KC = kF16 ? K / 2 : K;             // TMEM columns of one A image
    constexpr uint32_t kCols = (N + 2 * KC) <= 256 ? 256 : 512;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < N)
        for (int k = 0; k < K; ++k) {
            const float b = B[tid * K + k];
            if (kF16) {
                const __half bh = __float2half_rn(b), bl = __float2half_rn(b - __half2float(bh));
                *reinterpret_cast<__half*>(Bhi + canon_off16(tid, k, K)) = bh;
                *reinterpret_cast<__half*>(Blo + canon_off16(tid, k, K)) = bl;
            } else {
                const float bh = to_tf32(b), bl = to_tf32(b - bh);
                *reinterpret_cast<float*>(Bhi + canon_off(tid, k, K)) = bh;
                *reinterpret_cast<float*>(Blo + canon_off(tid, k, K)) = bl;
            }
        }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_hi = tmem + N, a_lo = tmem + N + KC;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < KC; c0 += 4) {
        uint32_t h[4], l[4];
        for (int j = 0; j < 4; ++j) {
            if (kF16) {
                const float a0 = A[tid * K + 2 * (c0 + j)], a1 = A[tid * K + 2 * (c0 + j) + 1];
                const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
                const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
                h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            } else {
                const float a = A[tid * K + c0 + j];
                const float ah = to_tf32(a);
                h[j] = __float_as_uint(ah); l[j] = __float_as_uint(to_tf32(a - ah));
            }
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a_hi + lane_base + c0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a_lo + lane_base + c0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // D = F32 (1 << 4); A / B format at bits 7 / 10: 0 = F16, 2 = TF32
    const uint32_t fmt = kF16 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int KS = kF16 ? K / 16 : K / 8;        // MMA k-steps: 16 fp16 or 8 tf32 = 32 bytes = two 16-byte chunks
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sbo = (K * EB / 16) * 128, lbo = 128;
        int first = 1;
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
            for (int ks = 0; ks < KS; ++ks) {
                const uint32_t koff = ks * 2 * 128;
                const uint64_t bh = make_desc(smem_u32(Bhi) + koff, lbo, sbo), bl = make_desc(smem_u32(Blo) + koff, lbo, sbo);
                auto mma = [&](uint32_t ta, uint64_t db) {
                    const uint32_t acc = first ? 0u : 1u;
                    if (kF16)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     :: "r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                     :: "r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    first = 0;
                };
                mma(a_lo + ks * 8, bh); mma(a_hi + ks * 8, bl); mma(a_hi + ks * 8, bh);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > 200000000LL) { if (tid == 0) *status = 1; break; }
        }
    }
    if (tid == 0) { t1 = clock64(); *cycles = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + lane_base + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kCols) : "memory");
}

template <int N, int K, bool kF16>
int run_rate(const char* name, float ascale) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(3);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f * ascale;     // ascale << 1: lo parts become fp16 subnormals
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
    float *dA, *dB, *dD; int* dS; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)N * K * (kF16 ? 2 : 4) * 2;
    cudaFuncSetAttribute(probe_ts_rate<N, K, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int rc = 0;
    for (int reps : {1, 65}) {
        cudaMemset(dS, 0, 4); cudaMemset(dD, 0, D.size() * 4);
        probe_ts_rate<N, K, kF16><<<1, 128, smem>>>(dA, dB, dD, dS, reps, dC);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; long long cyc = 0;
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        if (reps == 1) {
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double r = 0;
                    for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
                    maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
                    maxref = fmax(maxref, fabs(r));
                }
            printf("%s %s N=%d K=%d ascale=%g: cuda=%s timeout=%d max_abs_err=%.3e (max |ref| %.3g) rel=%.2e\n", name, kF16 ? "f16x2-split" : "tf32x3",
                   N, K, ascale, cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
        } else {
            printf("    %d repetitions of the 3-MMA GEMM: %lld cycles = %.0f per GEMM (cuda=%s timeout=%d)\n", reps, cyc, cyc / (double)reps, cudaGetErrorString(e), st);
        }
        rc |= e != cudaSuccess;
    }
    return rc;
}

int main() {
    if (run<128, 64, false>("gemm1")) return 1;
    if (run<128, 64, true>("gemm1")) return 1;
    if (run<64, 128, true>("gemm2")) return 1;
    if (run<64, 64, true>("node")) return 1;
    if (run<128, 64, true, true>("gemm1-ts")) return 1;
    if (run<64, 128, true, true>("gemm2-ts")) return 1;
    if (run_mn<128, 64, 128, false>("outer-mn")) return 1;
    if (run_mn<128, 64, 128, true>("outer-mn")) return 1;
    if (run_mn<128, 64, 128, false, 1>("outer-mn")) return 1;
    if (run_mn<128, 64, 128, true, 1>("outer-mn")) return 1;
    // round 2: fp16 two-way split against 3xTF32, accuracy and tensor-pipe rate (GEMM1: N=128 K=64; GEMM2: N=64 K=128)
    if (run_rate<128, 64, false>("gemm1-ts", 1.0f)) return 1;
    if (run_rate<128, 64, true>("gemm1-ts", 1.0f)) return 1;
    if (run_rate<128, 64, true>("gemm1-ts", 0.01f)) return 1;
    if (run_rate<64, 128, false>("gemm2-ts", 1.0f)) return 1;
    if (run_rate<64, 128, true>("gemm2-ts", 1.0f)) return 1;
    if (run_rate<64, 128, true>("gemm2-ts", 0.01f)) return 1;
    return 0;
}
