// neural_qc.cuh -- LDPCNeuralDecoder on the quasi-cyclic structure (5G BG2, Z = 32): no index tables at all.
//
// Same arithmetic, in the same order, as the generic one-kernel decoder (neural.cuh) and therefore as the per-layer
// chain CheckLayer -> VariableLayer -> ResidualLayer -> OutputLayer (models/layers.py:14-66, 78-125, 143-168,
// 180-210 composed as models/decoder.py documents) -- the outputs are bit-identical -- but the neighbour lists the
// reference materialises as (E,9) / (E,22) int64 tables are implied by the base graph:
//   * warp <-> one codeword, lane <-> circulant row; the other edges of a CHECK are the other cells of the base row,
//     reached by one __shfl_sync rotation each (lane + shift), the other edges of a VARIABLE are the other cells of
//     the base column, which live in the SAME lane once messages are kept variable-aligned;
//   * the edge-space state of a codeword -- c2v and a two-deep ring of variable outputs x (159 core cells x 3) --
//     lives in Tensor Memory (tcgen05.ld/st.32x32b: lane <-> TMEM lane, cell <-> column; 480 of the warp's 512
//     columns), the per-edge channel LLRs and the 38 degree-1 cells in shared memory, the per-edge weights w_ch
//     (25 KB, shared by the CTA's warps) in shared memory as [cell][lane];
//   * phase A (per base row): rotate x into check alignment, sign product / two smallest magnitudes with the
//     layer's zero rules, rotate the results back, store c2v;  phase B (per base column, lane-local): the sums
//     over the OTHER edges in the reference's order (running prefix + suffix chain), w_ch * llr + sum + residuals;
//   * the caller's edge numbering is create_LLR_mapping's (utils/ldpc_utils.py:62-95): variable-major,
//     e = 32*D_j + r*d_j + k for column j (D_j base edges before it, degree d_j), circulant row r, k-th check of
//     the variable.  Rows of llr_e / soft are moved with coalesced 128-byte accesses and transposed through a
//     pitch-33 shared-memory tile.
// The Python layer (models/decoder.py) takes this path only after checking that the caller's index tensors ARE
// those tables; anything else runs the table-driven kernels.
// One documented difference: a residual term that is not yet a queue entry (first one or two updates) is multiplied by a
// zero weight here instead of being skipped, so +-inf / NaN channel LLRs turn into NaN one update earlier than in the
// layer chain; finite inputs are unaffected (bit-identical).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace ldpc {

struct NeuralQcParams {
    const float* llr;      // [B, E] edge-space channel LLRs (variable-major edge order)
    const float* w_ch;     // [E]
    const float* w_res;    // [L]
    int L;                 // residual depth, 0..2
    int iters;
    long long B;
    const float* gt;       // [B, E] or null
    float* soft;           // [B, E]
    float* max_loss;       // [B] (with gt)
    // training forward: activations the backward pass reads, or null
    float* save_x;         // [iters, B, 197 cells, Z rows] input of every CheckLayer (x_0 = llr_e, x_1, ...), row-major per cell
    int32_t* argmax;       // [B] edge (caller's numbering) whose loss is the frame's maximum, or null
    // per-variable I/O (the trainer's call shape, training/trainer.py:95-110: LLRs and targets per code bit): llr and gt are
    // [B, N], every edge of a variable takes its variable's value (what llr[:, edge_to_var] would hold), soft is [B, N] and
    // holds the output at each variable's FIRST edge; max_loss / argmax still range over all E edges
    int per_var;
    float* star;           // [B, 2] (soft, target) at the arg-max edge, or null: all the backward needs of the two arrays
};

// backward of (max_loss) w.r.t. w_ch and w_res from the activations a save_x forward left behind
struct NeuralQcBwdParams {
    const float* save_x;   // [iters, B, 197, Z]
    const float* soft;     // [B, E]  (ignored when star is given)
    const float* gt;       // [B, E]
    const float* star;     // [B, 2] (soft, target) at the arg-max edge from a forward with `star`, or null
    const int32_t* argmax; // [B]
    const float* g_ml;     // [B] d(loss)/d(max_loss[b])
    const float* w_res;    // [L]
    int L;
    int iters;
    long long B;
    float* g_wch;          // [E]  accumulated (+=)
    float* g_wres;         // [2]  accumulated (+=), entries >= L stay untouched
};

// defined in neural_qc.cu (kernels: neural_qc_kernel.cuh)
int launch_neural_qc(const ldpc_code_t* c, const NeuralQcParams& p, cudaStream_t st);
int launch_neural_qc_bwd(const ldpc_code_t* c, const NeuralQcBwdParams& p, cudaStream_t st);

}  // namespace ldpc
