// gnn_node_pipe.cuh -- warp-specialised node kernel of the message-centred GNN decoder:
//   P[b][node][:] = W1B . mean_{e in node}(x[b][e][:] + emb[type(e)][:]) + b1
// (the aggregation half of MessageGNNLayer.forward, models/message_gnn_decoder.py:93-110: scatter-mean of the messages
// onto their variable / check node followed by the node's share of the first MLP layer).
//
// Why a second node kernel: gnn_node_tc_kernel (gnn_tc_pipe.cuh) runs gather -> split -> MMA -> epilogue -> store one
// after the other behind CTA barriers; ncu gave it 3.1 TB/s (the two launches per layer are 38 % of the forward) and the
// gather itself is a chain of dependent look-ups (node -> message list -> message type -> row) with ONE row in flight per
// thread for the degree-1 variable nodes that make up 73 % of a 5G base graph.  Here
//   * 16 loader warps do nothing but gather: a thread owns one 16-byte chunk of 4 adjacent nodes of the tile and walks
//     their (adjacent) message lists as one flat range, eight rows in flight whatever the degrees are; message index and
//     type come from ONE packed table entry, the type embeddings live in shared memory; the means land in a 3-deep ring
//     of staging tiles and the loaders never wait for the tensor pipe;
//   * 8 row warps (2 per TMEM lane quarter) take a finished tile: split hi/lo into Tensor Memory, one 3xTF32 GEMM
//     (tcgen05.mma, W1B hi/lo resident in shared memory), accumulator + b1 back into the same staging tile, coalesced
//     256-byte row stores, tile released.
// The sums are formed in the same order as in gnn_node_tc_kernel, so the two kernels are bit-identical.
// Shared memory: 32 KB weights + 3 x 32 KB staging + types x 256 B embeddings; Tensor Memory: 256 columns; 1 CTA per SM.
#pragma once
#include "gnn_tc_pipe.cuh"

namespace ldpc {

constexpr int kNpRowWarps = 8, kNpLoaderWarps = 16, kNpStages = 3;
constexpr int kNpRowThreads = kNpRowWarps * 32, kNpLoaderThreads = kNpLoaderWarps * 32;
constexpr int kNpThreads = kNpRowThreads + kNpLoaderThreads;                    // 768
constexpr int kNpListShift = 20;                                                // list entry = message | type << 20
constexpr int kNpMaxTypes = 256;
__host__ __device__ constexpr size_t node_pipe_smem(int types) {
    return (size_t)(2 * 64 * 64) * sizeof(float) + kNpStages * kPipeStage + (size_t)types * kH * sizeof(float);
}

__global__ void __launch_bounds__(kNpThreads, 1) gnn_node_pipe_kernel(
    const float* __restrict__ x, const float* __restrict__ emb_l, const float* __restrict__ packed_l, const float* __restrict__ tc_l,
    int kind, const int* __restrict__ ptr, const int* __restrict__ list2, int types, long long B, int E, int nodes,
    float* __restrict__ P, float* __restrict__ Msave, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* Whi = tc_smem;                                 // W1B [64 x 64]
    uint8_t* Wlo = Whi + 64 * 64 * 4;
    uint8_t* S0 = Wlo + 64 * 64 * 4;
    float* embs = reinterpret_cast<float*>(S0 + kNpStages * kPipeStage);
    __shared__ uint64_t full[kNpStages], freeb[kNpStages], mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float b1s[kH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        auto init = [&](uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory"); };
        for (int s = 0; s < kNpStages; ++s) { init(&full[s], kNpLoaderWarps); init(&freeb[s], kNpRowWarps); }
        init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(tc_l + (kind == 0 ? kTcW1BV : kTcW1BC));
        float4* dst = reinterpret_cast<float4*>(tc_smem);
        for (int t = tid; t < 2 * 64 * 64 / 4; t += kNpThreads) dst[t] = src[t];
        const float4* es = reinterpret_cast<const float4*>(emb_l);           // d_emb is a 16-byte aligned copy
        for (int t = tid; t < types * (kH / 4); t += kNpThreads) reinterpret_cast<float4*>(embs)[t] = es[t];
        if (tid < kH) b1s[tid] = packed_l[(kind == 0 ? kPkB1V : kPkB1C) + tid];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const long long rows = B * nodes, tiles = (rows + 127) / 128;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    bool ok = true;
    auto warp_arrive = [&](uint64_t* b) { __syncwarp(); if (lane == 0) mbar_arrive(b); };

    if (warp < kNpRowWarps) {
        // ================= row warps: thread <-> node row of the tile, 32 of the 64 columns =================
        const int rowi = tid & 127, part = tid >> 7;
        const uint32_t my_lane = ((uint32_t)((warp & 3) * 32)) << 16;
        const int cc = tid & 15, cr0 = tid >> 4;                              // cooperative mapping of the P store
        constexpr uint32_t kIdesc64 = umma_idesc_tf32(64);
        for (long long k = 0; k < my_tiles && ok; ++k) {
            const int s = (int)(k % kNpStages);
            const uint32_t use = (uint32_t)(k / kNpStages);
            uint8_t* S = S0 + (size_t)s * kPipeStage;
            const long long row0 = (blockIdx.x + k * gridDim.x) * 128;
            ok = mbar_wait(&full[s], use & 1u);
            if (!ok) break;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                const int col = part * 32 + c0;
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *stage_ptr(S, rowi, (col >> 2) + q);
                    v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
                }
                uint32_t hi[16], lo[16];
                split16(v, hi, lo);
                tmem_st16(tmem + my_lane + kTmNodeAHi + col, hi);
                tmem_st16(tmem + my_lane + kTmNodeALo + col, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, %0;" :: "n"(kNpRowThreads) : "memory");
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                umma_gemm3_ts(tmem + kTmNodeD, tmem + kTmNodeAHi, tmem + kTmNodeALo, smem_u32(Whi), smem_u32(Wlo), 64, 2048, kIdesc64);
                umma_commit(&mma_bar);
            }
            ok = mbar_wait(&mma_bar, (uint32_t)k & 1u);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                const int col = part * 32 + c0;
                float o[16];
                tmem_ld16(tmem + my_lane + kTmNodeD + col, o);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *stage_ptr(S, rowi, (col >> 2) + q) = make_float4(o[q * 4] + b1s[col + q * 4], o[q * 4 + 1] + b1s[col + q * 4 + 1],
                                                                      o[q * 4 + 2] + b1s[col + q * 4 + 2], o[q * 4 + 3] + b1s[col + q * 4 + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, %0;" :: "n"(kNpRowThreads) : "memory");
#pragma unroll
            for (int it = 0; it < 128 * 16 / kNpRowThreads; ++it) {
                const int rr = it * (kNpRowThreads / 16) + cr0;
                if (row0 + rr < rows) reinterpret_cast<float4*>(P + (size_t)(row0 + rr) * kH)[cc] = *stage_ptr(S, rr, cc);
            }
            warp_arrive(&freeb[s]);
        }
    } else {
        // ================= loader warps: gather-and-average, up to kNpStages tiles ahead of the row warps =================
        // 16 threads (one 16-byte chunk each) own FOUR ADJACENT nodes of the tile.  The message lists of adjacent nodes are
        // adjacent in the packed table, so the thread walks ONE flat range of entries, eight rows in flight, and adds each row
        // to the node it belongs to (list positions against the three inner boundaries).  nodes % 4 == 0 (host-checked), so a
        // group never straddles two codewords.
        const int ltid = tid - kNpRowThreads;
        const int cc = ltid & 15, nr = ltid >> 4;                             // chunk; node group in [0, 32): rows 4 nr .. 4 nr + 3
        const float4* embc = reinterpret_cast<const float4*>(embs) + cc;
        constexpr unsigned kEntMask = (1u << kNpListShift) - 1u;
        for (long long k = 0; k < my_tiles && ok; ++k) {
            const int s = (int)(k % kNpStages);
            const uint32_t use = (uint32_t)(k / kNpStages);
            uint8_t* S = S0 + (size_t)s * kPipeStage;
            const long long row0 = (blockIdx.x + k * gridDim.x) * 128;
            const long long r0 = row0 + 4 * nr;
            const bool valid = r0 < rows;
            int p[5] = {0, 0, 0, 0, 0};
            const float4* xb = reinterpret_cast<const float4*>(x) + cc;
            if (valid) {
                const long long b = r0 / nodes;
                const int node = (int)(r0 - b * nodes);
#pragma unroll
                for (int i = 0; i < 5; ++i) p[i] = __ldg(ptr + node + i);
                xb += (size_t)b * E * (kH / 4);
            }
            float4 m[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = p[0]; q < p[4]; q += 8) {
                unsigned ent[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) ent[u] = q + u < p[4] ? (unsigned)__ldg(list2 + q + u) : 0xffffffffu;
                float4 a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    a[u] = ent[u] != 0xffffffffu ? xb[(size_t)(ent[u] & kEntMask) * (kH / 4)] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (ent[u] != 0xffffffffu) {
                        const float4 em = embc[(ent[u] >> kNpListShift) * (kH / 4)];
                        const float4 v = make_float4(a[u].x + em.x, a[u].y + em.y, a[u].z + em.z, a[u].w + em.w);
                        const int qq = q + u;
                        if (qq < p[2]) {
                            if (qq < p[1]) { m[0].x += v.x; m[0].y += v.y; m[0].z += v.z; m[0].w += v.w; }
                            else { m[1].x += v.x; m[1].y += v.y; m[1].z += v.z; m[1].w += v.w; }
                        } else {
                            if (qq < p[3]) { m[2].x += v.x; m[2].y += v.y; m[2].z += v.z; m[2].w += v.w; }
                            else { m[3].x += v.x; m[3].y += v.y; m[3].z += v.z; m[3].w += v.w; }
                        }
                    }
            }
            ok = mbar_wait(&freeb[s], (use + 1u) & 1u);                       // the row warps have stored the tile that used this stage
            if (!ok) break;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = 4 * nr + i;
                const int d = p[i + 1] - p[i];
                const float inv = d > 0 ? 1.0f / (float)d : 0.0f;
                m[i].x *= inv; m[i].y *= inv; m[i].z *= inv; m[i].w *= inv;
                if (Msave && valid) reinterpret_cast<float4*>(Msave + (size_t)(r0 + i) * kH)[cc] = m[i];
                *stage_ptr(S, rr, cc) = m[i];
            }
            warp_arrive(&full[s]);
        }
    }
    if (!ok) { atomicExch(status, 1); asm volatile("trap;"); }               // a hand-off never arrived: fail loudly
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

}  // namespace ldpc
