/*
 * ldpc_b200.h -- C ABI of the B200-native LDPC message-passing engine.
 *
 * This is the drop-in boundary.  The reference project
 * (BananaFalls/LDPC-NeuralNetwork-Decoder) has no FFI layer: its boundary is the Python
 * class API.  Each entry point below states the reference interface it replaces
 * (paths relative to /root/reference/ldpc_neural_decoder).  The Python classes in
 * ldpc-neuralnetwork-decoder_b200/{models,utils} keep the reference signatures and call
 * these functions through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / ATen / pybind types.
 *  - Every function returns 0 on success or a negative ldpc_status code; the message of
 *    the last failure on the calling thread is returned by ldpc_last_error().
 *  - Device pointers are caller-owned (e.g. torch.Tensor.data_ptr()); the library never
 *    frees or retains them.  `stream` is a cudaStream_t passed as void*; launches are
 *    asynchronous on that stream unless the function name ends in _host.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails
 *    with LDPC_ERR_CUDA.
 *  - LLR sign convention is the reference's: bit = 1  <=>  belief < 0
 *    (models/traditional_decoders.py:99,252).
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1

typedef enum ldpc_status {
    LDPC_OK = 0,
    LDPC_ERR_INVALID = -1,     /* bad argument                                   */
    LDPC_ERR_UNSUPPORTED = -2, /* valid request the engine has no kernel for     */
    LDPC_ERR_CUDA = -3,        /* CUDA runtime error (message has the detail)    */
    LDPC_ERR_NOMEM = -4
} ldpc_status;

/* hard-decision output formats */
#define LDPC_HARD_F32 0    /* float32 0.0/1.0, [B,N]  (what the reference returns)        */
#define LDPC_HARD_U8 1     /* uint8 0/1, [B,N]                                            */
#define LDPC_HARD_PACKED 2 /* uint32 words, [B, ceil(N/32)], bit n%32 of word n/32        */

/* stopping rules */
#define LDPC_STOP_FIXED 0        /* run exactly `iters` iterations                          */
#define LDPC_STOP_PER_CODEWORD 1 /* freeze a codeword once its syndrome is zero             */
/* The reference's batch-global rule (stop when EVERY codeword of the batch is valid,
 * traditional_decoders.py:102-106,255-258) is built by the caller from `valid_mask`.     */

/* kernel selection */
#define LDPC_PATH_AUTO 0  /* fastest kernel that supports the code.  NOT reference operation order for
                           * min-sum: the specialised kernel forms v2c as posterior - own message, so soft
                           * outputs differ from the reference by rounding and +-inf channel LLRs can give
                           * inf - inf = NaN where the reference keeps inf.  Callers with non-finite LLRs
                           * (hard-decision inputs) must use LDPC_PATH_EXACT; the Python classes do that
                           * themselves.  BP under AUTO handles non-finite messages exactly.             */
#define LDPC_PATH_EXACT 1 /* table-driven kernel, reference operation order (any QC code)   */
#define LDPC_PATH_FAST 2  /* register/shuffle kernel specialised for a shipped 5G table     */

#define LDPC_ALGO_MINSUM 0
#define LDPC_ALGO_BP 1

typedef struct ldpc_code ldpc_code_t; /* opaque: base-graph shift tables resident on one device */

/* ---- library ---------------------------------------------------------------------- */
int ldpc_abi_version(void);
const char* ldpc_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench `gpu_launches`) */
uint64_t ldpc_launch_count(void);

/* ---- code tables -------------------------------------------------------------------
 * Replaces utils/ldpc_utils.py:97-125 (expand_base_matrix: dense H is never built) and the
 * adjacency scans traditional_decoders.py:26-40,161-175 / message_gnn_decoder.py:382-408.
 * shifts: [rows*cols] row-major, -1 = zero block, otherwise 0 <= shift < Z.
 * Lifting convention (ldpc_utils.py:121-123): check i*Z+r  <->  variable j*Z+((r+s) mod Z).
 * 1 <= Z <= 32.  Tables go to __constant__ memory when they fit a slot, else to global.
 * A lifting factor above 32 is passed as the equivalent code with Zs = a divisor of Z that is <= 32: a Z x Z circulant
 * with Z = m*Zs and shift s = m*q + t becomes, after renumbering r = m*a + b -> (b, a), the m blocks
 * (i*m + b, j*m + (b+t) mod m) with shift (q + (b+t) div m) mod Zs; the caller permutes LLRs / outputs accordingly
 * (utils/ldpc_utils.py: QCCode._split_lifting does both for the Python classes; output bit-identical to the natural code).
 * The table-driven decoder keeps a warp's messages in shared memory when enough warps fit, else in a stream-ordered
 * global workspace (cudaMallocAsync on the caller's stream), so code size is bounded by node degree <= 32 only.        */
int ldpc_code_create(const int16_t* shifts, int rows, int cols, int Z, int device, ldpc_code_t** out);
int ldpc_code_destroy(ldpc_code_t* code);
/* info[0..7] = rows, cols, Z, base edges, N=cols*Z, M=rows*Z, max row degree, max col degree */
int ldpc_code_info(const ldpc_code_t* code, int32_t info[8]);
/* 1 if the specialised (LDPC_PATH_FAST) kernels cover this code/algorithm */
int ldpc_code_has_fast_path(const ldpc_code_t* code, int algo);

/* ---- classic decoders --------------------------------------------------------------
 * ldpc_minsum_decode replaces MinSumScaledDecoder.decode (traditional_decoders.py:177-260),
 * ldpc_bp_decode replaces BeliefPropagationDecoder.decode (traditional_decoders.py:42-109).
 * Flooding schedule, fp32, no clipping (BP keeps the reference's inf/NaN behaviour).
 *   llr         [B,N] fp32 row-major, device
 *   soft_out    [B,N] fp32 posterior beliefs (reference local `var_beliefs`) or NULL
 *   hard_out    per hard_dtype, or NULL
 *   syndrome_ok [B] uint8: 1 if the returned hard decision satisfies every check
 *               (_check_valid_codeword, traditional_decoders.py:111-134,262-284) or NULL
 *   iters_out   [B] int32 iterations actually run per codeword, or NULL
 *   valid_mask  [B,mask_words] uint64, bit t of word t/64 = "valid after iteration t+1",
 *               or NULL (mask_words ignored)                                            */
int ldpc_minsum_decode(const ldpc_code_t* code, const float* llr, int64_t B, int iters, float alpha,
                       int stop_mode, int path, float* soft_out, void* hard_out, int hard_dtype,
                       uint8_t* syndrome_ok, int32_t* iters_out, uint64_t* valid_mask, int mask_words,
                       void* stream);
int ldpc_bp_decode(const ldpc_code_t* code, const float* llr, int64_t B, int iters, int stop_mode, int path,
                   float* soft_out, void* hard_out, int hard_dtype, uint8_t* syndrome_ok, int32_t* iters_out,
                   uint64_t* valid_mask, int mask_words, void* stream);

/* flag[0] |= 1 if any of x[0..n) is +-inf or NaN (x 16-byte aligned; flag is a device int the caller zeroes).  One read-only
 * pass; the drop-in classes use it to route non-finite LLR batches to LDPC_PATH_EXACT (see LDPC_PATH_AUTO above).            */
int ldpc_nonfinite_flag(const float* x, int64_t n, int32_t* flag, void* stream);

/* ldpc_syndrome_check replaces _check_valid_codeword (traditional_decoders.py:111-134,262-284):
 * syndrome_ok[b] = 1 iff every parity check of H is satisfied by hard[b] (per hard_dtype).    */
int ldpc_syndrome_check(const ldpc_code_t* code, const void* hard, int hard_dtype, int64_t B, uint8_t* syndrome_ok,
                        void* stream);

/* ldpc_encode: systematic encoder, absent from the reference (every loop there sends the all-zero codeword,
 * trainer.py:86, comparative_evaluation.py:132).  info: [B, K] uint8 0/1 (device), codeword:
 * [B, N] uint8 0/1 with codeword[:, :K] == info and every parity check of the code satisfied.  plan / binv: the device
 * tables built by utils/encoder.py for a code of the shape H = [A B 0; C D I] (plan = [g, kb, words, core_rows[g],
 * core_cols[g], ext_of_row[rows]] int32, binv = B^-1 packed by rows, [g*Z][words] uint32).                          */
int ldpc_encode(const ldpc_code_t* code, const uint8_t* info, int64_t B, const int32_t* plan, int64_t plan_len,
                const uint32_t* binv, uint8_t* codeword, void* stream);

/* 5G NR rate matching (3GPP TS 38.212 5.4.2.1 bit selection + 5.4.2.2 bit interleaving) and the receiver-side inverse
 * that produces the punctured / rate-matched LLR layout the decoders consume.  No counterpart in the reference, which
 * feeds its decoders the full N-long, un-punctured all-zero codeword (trainer.py:86, comparative_evaluation.py:132,
 * utils/channel.py:205-231); SURVEY.md section 8 f3.  Tables come from utils/rate_match.py (device int32 / fp32):
 *   sel     [E]    codeword position (0..N-1) carried by transmitted bit t
 *   inv_ptr [N+1], inv_idx [E]   CSR inverse of sel (transmissions of position n, in circular-buffer order)
 *   base    [N]    LLR a position starts from: 0 (punctured first 2Z columns, untransmitted tail), the known-zero
 *                  LLR for filler bits
 * ldpc_rate_match:   out[b][t] = codeword[b][sel[t]]                       (uint8 0/1)
 * ldpc_rate_recover: llr_out[b][n] = base[n] + sum_k rx_llr[b][inv_idx[k]]   (soft-combines repetitions)        */
int ldpc_rate_match(const uint8_t* codeword, const int32_t* sel, int64_t B, int64_t N, int64_t E, uint8_t* out, void* stream);
int ldpc_rate_recover(const float* rx_llr, const int32_t* inv_ptr, const int32_t* inv_idx, const float* base, int64_t B,
                      int64_t N, int64_t E, float* llr_out, void* stream);

/* Host-buffer variant (the `e2e` path of bench.py): llr_host / hard_host are HOST pointers
 * (pinned memory recommended).  The call chunks the batch, overlaps H2D, decode and D2H on
 * internal streams and returns when hard_host (and soft_host if given) are complete.     */
int ldpc_decode_host(const ldpc_code_t* code, int algo, const float* llr_host, int64_t B, int iters, float alpha,
                     int path, float* soft_host, void* hard_host, int hard_dtype, int64_t chunk);

/* Quantised host LLRs.  A 5G receiver's demapper produces 8-bit (or half-precision) LLRs; shipping them as such
 * cuts the PCIe bytes per codeword 4x (2x) -- the host path is PCIe-bound, see DESIGN.md section 4.  The decoder
 * still computes in fp32 on llr = (float)q * llr_scale (LDPC_LLR_I8) or llr = (float)h * llr_scale (LDPC_LLR_F16);
 * the outputs are bit-identical to the reference decoders (traditional_decoders.py:37-109,178-260) run on those
 * float values.  LDPC_LLR_F32 ignores llr_scale and is ldpc_decode_host.                                       */
enum { LDPC_LLR_F32 = 0, LDPC_LLR_F16 = 1, LDPC_LLR_I8 = 2 };
int ldpc_decode_host_q(const ldpc_code_t* code, int algo, const void* llr_host, int llr_format, float llr_scale,
                       int64_t B, int iters, float alpha, int path, float* soft_host, void* hard_host,
                       int hard_dtype, int64_t chunk);

/* ---- channel + metrics -------------------------------------------------------------
 * ldpc_awgn_llr replaces AWGNChannel.transmit (utils/channel.py:205-231): BPSK 0->+1,
 * sigma = 1/sqrt(10^(snr_db/10)), llr = 2*(s+n)/sigma^2, noise from Philox4x32-10 keyed by
 * (seed) with counter (first_frame + b, n/4): results do not depend on batch split or GPU
 * count.  bits: [B,N] uint8 or NULL (= all-zero codeword, as every reference sweep uses). */
int ldpc_awgn_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, uint64_t seed, uint64_t first_frame,
                  float* llr_out, void* stream);

/* ldpc_qpsk_llr replaces the chain qpsk_modulate -> awgn_channel -> qpsk_demodulate (utils/channel.py:4-60, 62-89,
 * 91-154) that every shipped training / evaluation loop of the reference uses (trainer.py:89-95,
 * comparative_evaluation.py:138-146): bit 2k rides on I and bit 2k+1 on Q of symbol k with amplitude 1/sqrt(2), each
 * component sees N(0, 1/(2*snr_linear)), llr = 2*r/noise_var with noise_var = 1/snr_linear -- the reference's
 * scaling (its LLRs are 1/sqrt(2) of the true ones; min-sum is scale-invariant, BP and the GNN are not).
 * true_llr = 1 selects the exact LLR 2*sqrt(2)*r*snr_linear instead.  Noise as in ldpc_awgn_llr.  [B,N] out.   */
int ldpc_qpsk_llr(const uint8_t* bits, int64_t B, int64_t N, float snr_db, int true_llr, uint64_t seed,
                  uint64_t first_frame, float* llr_out, void* stream);
/* ldpc_count_errors replaces compute_ber_fer (utils/channel.py:156-190) with integer
 * counters: counters[0]+=bit errors, [1]+=frame errors, [2]+=frames.  tx may be NULL
 * (all-zero).  hard per hard_dtype.                                                      */
int ldpc_count_errors(const void* hard, int hard_dtype, const uint8_t* tx, int64_t B, int64_t N,
                      uint64_t* counters, void* stream);
/* Fused Monte-Carlo path (SURVEY K4+K1/K2+K5): generate LLRs on chip, decode, count; LLRs
 * never touch HBM.  counters (device, 4 x uint64) += [bit errors, frame errors, frames,
 * undetected frame errors (syndrome ok but bits wrong)].                                 */
int ldpc_sim_fer(const ldpc_code_t* code, int algo, int iters, float alpha, float snr_db, uint64_t seed,
                 uint64_t first_frame, uint64_t n_frames, uint64_t* counters, void* stream);

/* ---- edge-space layers (models/layers.py) ------------------------------------------
 * x, out: [B,E] fp32.  idx: [E,K] int64 neighbour lists, -1 padded
 * (utils/ldpc_utils.py:5-60).                                                            */
/* CheckLayer.forward, layers.py:14-66.  argmin_out [B,E] int32 (slot of the selected
 * minimum, -1 if none) is optional and feeds the backward.                               */
int ldpc_check_layer_fwd(const float* x, const int64_t* idx, int64_t B, int64_t E, int K, float* out,
                         int32_t* argmin_out, void* stream);
int ldpc_check_layer_bwd(const float* x, const int64_t* idx, const int32_t* argmin, const float* grad_out,
                         int64_t B, int64_t E, int K, float* grad_x, void* stream);
/* VariableLayer.forward, layers.py:78-125: out = llr + sum_k c2v[idx]                    */
int ldpc_variable_layer_fwd(const float* llr, const float* c2v, const int64_t* idx, int64_t B, int64_t E, int K,
                            float* out, void* stream);
/* grad_c2v[b, idx[e,k]] += grad_out[b,e]  (grad_llr = grad_out, done by the caller)      */
int ldpc_variable_layer_bwd(const int64_t* idx, const float* grad_out, int64_t B, int64_t E, int K,
                            float* grad_c2v, void* stream);
/* ResidualLayer.forward, layers.py:143-168: out = w_ch*llr + c2v + sum_{i<L} w_res[i]*prev[i]
 * prev: array of L device pointers (host array of pointers), each [B,E].                 */
int ldpc_residual_layer_fwd(const float* llr, const float* c2v, const float* w_ch, const float* w_res,
                            const float* const* prev, int L, int64_t B, int64_t E, float* out, void* stream);
/* Variable + residual update of the unrolled neural min-sum decoder (LDPCNeuralDecoder; the
 * reference's models/decoder.py is missing, prototype: EE4002R_2025.ipynb cell 11
 * `variable_layer_update`): out = w_ch*llr + sum_k c2v[idx] + sum_{i<L} w_res[i]*prev[i],
 * bit-identical to VariableLayer (layers.py:78-125, zero llr) then ResidualLayer (:143-168). */
int ldpc_neural_variable_layer_fwd(const float* llr, const float* c2v, const int64_t* idx, const float* w_ch,
                                   const float* w_res, const float* const* prev, int L, int64_t B, int64_t E, int K,
                                   float* out, void* stream);
/* The gather-type layers on "sorted-pack" tables (same arithmetic and results as the int64
 * entry points above; layers.py:14-66, :78-125, notebook cell 11): idx16 = [K,E] uint16 from
 * ldpc_neural_pack_index on the table with every row compacted (valid entries first, caller's
 * order); cnt [E] uint8 = valid entries per column; perm [E] uint16 (or NULL) = edge of column
 * t, columns ordered by descending cnt so that a warp skips the padding.
 * check: nstar (optional) [B,E] int32 = edge selected as the minimum, -1 if none; its backward
 * needs only (x, out, nstar).  variable: w_ch == NULL -> out = llr + sum (VariableLayer);
 * else out = w_ch*llr + sum + sum_{i<L} w_res[i]*prev[i] (variable + residual update).      */
int ldpc_check_layer_fwd_sorted(const float* x, const uint16_t* idx16, int K, const uint8_t* cnt, const uint16_t* perm,
                                int64_t B, int64_t E, float* out, int32_t* nstar, void* stream);
int ldpc_variable_layer_fwd_sorted(const float* llr, const float* c2v, const uint16_t* idx16, int K, const uint8_t* cnt,
                                   const uint16_t* perm, const float* w_ch, const float* w_res, const float* const* prev,
                                   int L, int64_t B, int64_t E, float* out, void* stream);
int ldpc_check_layer_bwd_nstar(const float* x, const float* out, const int32_t* nstar, const float* grad_out, int64_t B,
                               int64_t E, float* grad_x, void* stream);
/* Whole LDPCNeuralDecoder forward (inference / validation) in one kernel: `iters` unrolled
 * iterations of CheckLayer (layers.py:14-66) -> VariableLayer (:78-125) -> ResidualLayer
 * (:143-168), the last check messages summed per variable into OutputLayer (:180-210); the
 * composition of the reference's missing models/decoder.py (notebook cell 11 `forward`).
 * Messages stay in shared memory across iterations; bit-identical to the per-layer entry
 * points.  cidx/vidx are the [E,K] int64 neighbour tables packed ONCE per code by
 * ldpc_neural_pack_index to [K,E] uint16, with every row COMPACTED first (valid entries
 * first, caller's order; unused slots any index < E); ccnt / vcnt [E] uint8 = number of
 * valid entries per column; cperm / vperm (optional, [E] uint16): column t belongs to edge
 * perm[t] -- the host sorts columns by descending count so that warps skip the padding;
 * NULL = identity.  L <= 4.  gt_e / max_loss optional.                                   */
int ldpc_neural_pack_index(const int64_t* idx, int64_t E, int K, uint16_t* out, void* stream);
int ldpc_neural_decode(const float* llr_e, const uint16_t* cidx, int Kc, const uint8_t* ccnt, const uint16_t* cperm,
                       const uint16_t* vidx, int Kv, const uint8_t* vcnt, const uint16_t* vperm, const float* w_ch,
                       const float* w_res, int L, int iters, int64_t B, int64_t E, const float* gt_e, float* soft,
                       float* max_loss, void* stream);
/* The same decoder on the quasi-cyclic structure of the code (csrc/neural_qc.cuh): no index tables -- the neighbour
 * lists of create_LLR_mapping (utils/ldpc_utils.py:62-95) are implied by the base graph, check neighbours are lane
 * rotations, variable neighbours are lane-local, the edge state lives in Tensor Memory.  Compiled for the 5G BG2 tables
 * at Z = 32, 16 (the reference's default --lifting_factor), 8 and 4 (32 / Z codewords per warp) -- LDPC_ERR_UNSUPPORTED
 * otherwise -- and residual depth L <= 2.  llr_e / gt_e / soft: [B, E] in the variable-major
 * edge order of create_LLR_mapping(H.T); bit-identical to ldpc_neural_decode with that code's tables.
 * Training (the loop of training/trainer.py:95-110, `loss.mean().backward()`): pass save_x [iters, B, 197, Z] (the input of
 * every CheckLayer, stored lane-major) and argmax [B] (edge whose BCE is the frame's max_loss); ldpc_neural_backward_qc then
 * accumulates d(sum_b g_ml[b] * max_loss[b]) / d w_ch into g_wch [E] and / d w_res into g_wres [L] (+=) -- the autograd of
 * the four reference layers in that composition, one forward and one backward kernel per step.                        */
int ldpc_neural_decode_qc(const ldpc_code_t* code, const float* llr_e, const float* w_ch, const float* w_res, int L, int iters,
                          int64_t B, const float* gt_e, float* soft, float* max_loss, float* save_x, int32_t* argmax, void* stream);
int ldpc_neural_backward_qc(const ldpc_code_t* code, const float* save_x, const float* soft, const float* gt_e, const int32_t* argmax,
                            const float* g_ml, const float* w_res, int L, int iters, int64_t B, float* g_wch, float* g_wres,
                            void* stream);
/* The same two kernels with PER-VARIABLE input and output -- the trainer's own call shape (training/trainer.py:95-110,180-187:
 * LLRs and targets per code bit, expanded with llr[:, edge_to_var] before the layers run): llr_v / gt_v / soft_v are [B, N]
 * (N = 52 Z); every edge of a variable takes its variable's LLR and target, soft_v holds the output at each variable's
 * first edge, max_loss / argmax range over all E edges exactly as in ldpc_neural_decode_qc (bit-identical to it on the
 * expanded arrays).  star [B, 2] receives (soft, target) at the arg-max edge, which is all ldpc_neural_backward_qc_var
 * needs of the two arrays: no [B, E] tensor exists anywhere on this path.                                             */
int ldpc_neural_decode_qc_var(const ldpc_code_t* code, const float* llr_v, const float* w_ch, const float* w_res, int L, int iters,
                              int64_t B, const float* gt_v, float* soft_v, float* max_loss, float* save_x, int32_t* argmax,
                              float* star, void* stream);
int ldpc_neural_backward_qc_var(const ldpc_code_t* code, const float* save_x, const float* star, const int32_t* argmax,
                                const float* g_ml, const float* w_res, int L, int iters, int64_t B, float* g_wch, float* g_wres,
                                void* stream);
/* OutputLayer.forward, layers.py:180-210: soft = sigmoid(final+llr); if gt: per-row max of
 * BCE(soft, gt) -> max_loss [B], argmax [B] int32 (for the backward).                    */
int ldpc_output_layer_fwd(const float* final_llr, const float* llr, const float* gt, int64_t B, int64_t E,
                          float* soft, float* max_loss, int32_t* argmax, void* stream);

/* ---- message-centred GNN (models/message_gnn_decoder.py) ---------------------------
 * Weights are the reference nn.Module's parameters, flattened by the Python wrapper into
 * one fp32 buffer in state_dict order per layer (see models/message_gnn_decoder.py in the
 * package).  The graph is the code's Tanner graph, messages in check-major /
 * ascending-variable order (message_gnn_decoder.py:397-406); the dense E x E adjacencies
 * (:423-467) are replaced by per-variable / per-check segment means.                     */
typedef struct ldpc_gnn ldpc_gnn_t;
int ldpc_gnn_create(const ldpc_code_t* code, int num_layers, int hidden, int num_types, const int32_t* edge_type /*[base edges]*/,
                    ldpc_gnn_t** out);
int ldpc_gnn_destroy(ldpc_gnn_t* g);
size_t ldpc_gnn_param_count(const ldpc_gnn_t* g);
size_t ldpc_gnn_workspace_bytes(const ldpc_gnn_t* g, int64_t B, int training);
/* MessageGNNDecoder.forward (message_gnn_decoder.py:190-317): soft_out = combined LLR
 * (pre-sigmoid), prob_out = sigmoid(soft_out).  Either may be NULL.                       */
int ldpc_gnn_forward(const ldpc_gnn_t* g, const float* params, const float* llr, int64_t B, float* soft_out,
                     float* prob_out, void* workspace, size_t ws_bytes, int training, void* stream);
/* mean-BCE loss on prob_out vs gt (message_gnn_decoder.py:313-315) and its gradient w.r.t.
 * every parameter (autograd of the reference module); needs the workspace of a
 * training=1 forward with the SAME params.  grad_params is accumulated into (+=).  The weight
 * images are re-packed from `params` here, so other forwards on the handle in between are
 * harmless; a handle is single-stream (its packed-weight scratch is shared by all calls).  */
int ldpc_gnn_backward(const ldpc_gnn_t* g, const float* params, const float* llr, const float* gt, int64_t B,
                      float* loss_out, float* grad_params, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
