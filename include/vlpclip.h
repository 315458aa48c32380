/* vlpclip.h -- C ABI of the B200-native fused CLIP (symmetric InfoNCE) head.
 *
 * Drop-in boundary for the hot path of
 *   /root/reference/src/models/pretrain/VisionLanguageModule.py
 *     :448-453  projection + L2-normalise of both streams      -> vlpclip_project_normalize_*
 *     :456-459  clamp(exp(logit_scale)) scaled N x N similarity -> fused into vlpclip_lse_fwd / vlpclip_grad
 *     :550-552  row / column cross-entropy                      -> vlpclip_lse_fwd + vlpclip_lse_merge + vlpclip_loss_reduce
 *     autograd of the above (triggered after :645)              -> vlpclip_grad, vlpclip_normalize_bwd
 *
 * The reference has no FFI of its own (pure Python over torch ops); the Python host side binds
 * these symbols with ctypes. Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocator); the library
 *     never allocates persistent device memory and never frees -- the one exception is the peer
 *     window of the sharded backward (vlpclip_peer_*), which must be exportable over CUDA IPC;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = ok, negative = error; vlpclip_last_error() returns a thread-local message;
 *   - the N x N logit matrix is never written to global memory by any entry point.
 *
 * `scale` is always a DEVICE pointer to one fp32 value s = clamp(exp(logit_scale), max=100)
 * (VisionLanguageModule.py:456-457): the host never needs the temperature, so a training step
 * issues no host<->device synchronisation and can be captured in a CUDA graph.
 * "log2 domain": running maxima `m` and sums `l` describe  sum_j exp(S_ij) = l * 2^m .
 */
#ifndef VLPCLIP_H_
#define VLPCLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLPCLIP_VERSION 100

int vlpclip_version(void);
const char* vlpclip_last_error(void);

/* number of SMs of the current device (grid sizing / workspace sizing) */
int vlpclip_sm_count(void);

/* cap the number of SMs the persistent kernels occupy (0 = all); returns the effective count.
 * Used by the sharded variant to leave a few SMs to overlapping NCCL kernels. */
int vlpclip_set_sm_limit(int n_sms);

/* kernels launched by this library so far in this process (bench.py: gpu_launches) */
unsigned long long vlpclip_launch_count(void);

/* bf16 -> fp16 copy (values of unit-norm embeddings are exactly representable but for |x| < 2^-14) */
int vlpclip_cast_bf16_to_f16(const void* src_bf16, void* dst_f16, size_t n_elems, void* stream);

/* fp32 embeddings -> bf16 operand copy and (dst_f16 != NULL) the fp16 image of the bf16-rounded values,
 * one pass (n_elems multiple of 4): the operands of the forward sweep / the backward GEMMs */
int vlpclip_cast_f32_operands(const float* src_f32, void* dst_bf16, void* dst_f16, size_t n_elems,
                              void* stream);

/* ---- forward: per-row log-sum-exp statistics of  S = scale * X Y^T  ----
 * X: [n_rows, d] bf16 row-major (row stride ldx elements), Y: [n_cols, d] bf16 (row stride ldy).
 * The positive pair of row i is column i - diag_shift (if inside [0, n_cols)).
 * Outputs (each [n_rows] fp32):
 *   row_max[i] = max_j <X_i, Y_j>                       (raw cosine, NOT scaled, positive included)
 *   row_l[i]   = sum_{j != positive} exp(scale*<X_i,Y_j>) / 2^fl(k*row_max[i]),  k = scale*log2(e)
 *   diag[i]    = <X_i, Y_{i - diag_shift}>  (rows without a partner column are left untouched)
 * (max, l) pairs from different column ranges / ranks are combined with vlpclip_lse_merge.
 * workspace: vlpclip_lse_workspace_bytes(n_rows, n_cols, d) bytes.
 * Replaces VisionLanguageModule.py:459 + the log-sum-exp half of :550 (rows) / :551 (columns,
 * by calling it with X and Y swapped).
 */
size_t vlpclip_lse_workspace_bytes(int n_rows, int n_cols, int d);
int vlpclip_lse_fwd(const void* x_bf16, int ldx, const void* y_bf16, int ldy, int n_rows,
                    int n_cols, int d, const float* scale, int diag_shift, float* row_max,
                    float* row_l, float* diag, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Same sweep, but ALSO the statistics of the columns of S over the given rows (one pass over
 * the logits instead of two): col_max[j] / col_l[j] ([n_cols] fp32) follow the (max, l) convention
 * of the rows, except that col_max is an upper reference (>= the true column maximum) rather than
 * the maximum itself; the positive pair (row j + diag_shift) is left out of col_l as well.
 * Column partial sums are carried with 2^100 headroom: a term is dropped only if it lies more than
 * 226 log2 units (157 nats) below the largest row maximum of its 32-row group.
 * Replaces VisionLanguageModule.py:459 + the log-sum-exp halves of :550 AND :551. */
size_t vlpclip_lse_fused_workspace_bytes(int n_rows, int n_cols, int d);
int vlpclip_lse_fwd_fused(const void* x_bf16, int ldx, const void* y_bf16, int ldy, int n_rows,
                          int n_cols, int d, const float* scale, int diag_shift, float* row_max,
                          float* row_l, float* diag, float* col_max, float* col_l, void* workspace,
                          size_t workspace_bytes, void* stream);

/* merge `nparts` partial (max, l) pairs laid out [nparts][n] (fixed order) and fold the positive
 * pair logits `diag` [n] (may be NULL = no positive pair) back in.  Any output may be NULL:
 *   lse      natural-log LSE of the full row
 *   out_max / out_l   merged pair
 *   out_lg2l log2(sum_j exp(S_ij)) - k*max   (the form the backward consumes)
 *   out_q    1 - softmax probability of the positive pair
 *   out_loss lse - scale*diag  (per-row cross-entropy, log1p-accurate when the positive dominates;
 *            requires diag) */
int vlpclip_lse_merge(const float* part_max, const float* part_l, const float* diag, int nparts,
                      int n, const float* scale, float* lse, float* out_max, float* out_l,
                      float* out_lg2l, float* out_q, float* out_loss, void* stream);

/* out2[0] = sum_i row_loss[i], out2[1] = sum_i col_loss[i] over n entries (either may be NULL);
 * single block, fixed order (reproducible). The caller divides by the global batch size
 * (VisionLanguageModule.py:550-552). */
int vlpclip_loss_reduce(const float* row_loss, const float* col_loss, int n, float* out2,
                        void* stream);

/* ---- backward: dX = scale * G Y,
 *      G = (w_row P_row + w_col P_col - (w_row + w_col) delta) / (2 n_global) ----
 * (w_row = w_col = 1 is the symmetric loss of :552; other non-negative weights serve callers
 *  that back-propagate image_loss / text_loss separately.)
 * X, Y: fp16 copies of the embeddings ([n_rows, d] / [n_cols, d], row strides ldx / ldy).
 * (x_max, x_lg2l, x_q)[n_rows], (y_max, y_lg2l, y_q)[n_cols]: merged statistics
 * (vlpclip_lse_merge) of the rows of S owned by X / by Y.  delta_ij = 1 iff i == j + diag_shift.
 * dX: [n_rows, d] (row stride d), overwritten; fp32, or bf16 when dx_bf16 != 0.
 * out_mul (optional device scalar, may be NULL): multiplied into dX (the upstream gradient of
 * loss.backward(); dscale is NOT multiplied).
 * dscale (optional, may be NULL): receives sum_ij G_ij <X_i, Y_j> (fp32, one value).
 * Replaces autograd of VisionLanguageModule.py:459, :550-552.
 */
size_t vlpclip_grad_workspace_bytes(int n_rows, int n_cols, int d);
int vlpclip_grad(const void* x_f16, int ldx, const void* y_f16, int ldy, const float* x_max,
                 const float* x_lg2l, const float* x_q, const float* y_max, const float* y_lg2l,
                 const float* y_q, int n_rows, int n_cols, int d, const float* scale, int diag_shift,
                 int n_global, float w_row, float w_col, const float* out_mul, int dx_bf16, void* dx,
                 float* dscale, void* workspace, size_t workspace_bytes, void* stream);

/* the same two sweeps on fp16 operands (the backward's operand copies; bf16-representable inputs
 * convert exactly, so the statistics are identical): lets a sharded step gather ONE fp16 copy of T */
int vlpclip_lse_fwd_f16(const void* x_f16, int ldx, const void* y_f16, int ldy, int n_rows, int n_cols,
                        int d, const float* scale, int diag_shift, float* row_max, float* row_l,
                        float* diag, void* workspace, size_t workspace_bytes, void* stream);
int vlpclip_lse_fwd_fused_f16(const void* x_f16, int ldx, const void* y_f16, int ldy, int n_rows,
                              int n_cols, int d, const float* scale, int diag_shift, float* row_max,
                              float* row_l, float* diag, float* col_max, float* col_l,
                              void* workspace, size_t workspace_bytes, void* stream);
/* bf16 -> fp16 cast of n_elems (multiple of 8) values stored to n_dst (<= 8) destinations at once.
 * dsts is a HOST array of device pointers; with dsts[r] inside rank r's peer window this is the
 * all-gather of the local text shard, fused into the cast that the backward needs anyway. */
int vlpclip_cast_push_f16(const void* src_bf16, size_t n_elems, void* const* dsts, int n_dst,
                          void* stream);

/* s = min(exp(logit_scale), 100) and ds/dlogit_scale on the device (VisionLanguageModule.py:456-457);
 * logit_scale is one fp32 (is_f64 = 0) or fp64 (is_f64 = 1, the reference's dtype) device value. */
int vlpclip_scale_prep(const void* logit_scale, int is_f64, float* scale, float* dscale_dls,
                       void* stream);
/* out3 = {loss, image_loss, text_loss} from the two (all-reduced) loss sums (:550-552) */
int vlpclip_loss_finish(const float* sums2, int n_global, float* out3, void* stream);

/* measurement hook: bracket the grad_pair_kernel launches of subsequent vlpclip_grad calls with
 * CUDA events on their stream (enable = 1) and read the duration of the latest call in ms */
int vlpclip_time_grad_kernel(int enable);
float vlpclip_last_grad_kernel_ms(void);
/* development builds only (-DVLP_PROFILE_WAITS; tools/check_grad_both.py --prof, tools/wait_profile.py):
 * device buffer of [SMs][16] int64 in which the next backward kernels record, per role, the cycles
 * spent blocked on each pipeline barrier; a no-op in the shipped library */
int vlpclip_dev_set_wait_profile(void* buf);

/* host-only: the backward's work partition for n_clusters SM pairs (see grad_bwd.cu, "stream-K").
 * seg rows = {cluster, row block, t0, t1, slot} (slot -1: whole row block, final rows written by the
 * kernel; 0 / 1: partial block of that cluster), red rows = {row block, cluster, slot} in the
 * order the pieces of split row blocks are summed.  Used by the CPU test-suite. */
int vlpclip_grad_plan(int n_row_blocks, int tiles, int n_clusters, int* seg, int max_seg, int* n_seg,
                      int* red, int max_red, int* n_red);

/* ---- sharded backward: dX rows go straight to the rank that owns them (fused reduce-scatter) ----
 * Same computation as vlpclip_grad, but final row r is stored (fp32, no out_mul) at
 *   owner_rows[r / rows_per_owner] + (r % rows_per_owner) * d
 * where owner_rows[o] points into rank o's peer window (NVLink peer memory, mapped with
 * vlpclip_peer_open) at the slot reserved for the calling rank; the stores leave the SM from the
 * accumulator epilogue, so the transfer overlaps the remaining tiles. After a cross-rank barrier
 * every rank sums the slots of its own window in slot order with vlpclip_slot_sum (deterministic).
 * Replaces reduce_scatter(dT_all) of the global-batch loss; owner_rows is a
 * HOST array of n_owners (<= 8) device pointers.
 */
int vlpclip_grad_scatter(const void* x_f16, int ldx, const void* y_f16, int ldy, const float* x_max,
                         const float* x_lg2l, const float* x_q, const float* y_max,
                         const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                         const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                         void* const* owner_rows, int n_owners, int rows_per_owner, float* dscale,
                         void* workspace, size_t workspace_bytes, void* stream);
/* ---- single-recompute backward: dX, dY and dscale from ONE sweep over the logit tiles ----
 * Same mathematics as two vlpclip_grad calls ((X, Y) and (Y, X)), but every tile of
 * G = d loss / d S is formed once and feeds both accumulations (8 N^2 D executed flops per training
 * step instead of 10 N^2 D):
 *   dX[n_rows, d] = scale * G   Y   (rows of X; fp32 or bf16, multiplied by out_mul if given)
 *   dY[n_cols, d] = scale * G^T X   (rows of Y; fp32 or bf16, multiplied by out_mul if given), or,
 *                   when dy_owner_rows != NULL, stored fp32 without out_mul into the owning ranks'
 *                   peer windows exactly like vlpclip_grad_scatter (fused reduce-scatter; dy ignored)
 *   dscale        = sum_ij G_ij <X_i, Y_j>  (optional)
 * Statistics, weights, diag_shift and n_global as for vlpclip_grad with (X, Y) in this order.
 * Runs as ONE persistent kernel on all SMs (producer / dI-consumer SM pairs + dT-consumer SMs that
 * exchange G tiles through an L2-resident ring inside the workspace); the sums are formed in a
 * fixed order, so results are bit-reproducible.
 * Replaces autograd of VisionLanguageModule.py:459, :550-552 (both embedding gradients at once). */
size_t vlpclip_grad_both_workspace_bytes(int n_rows, int n_cols, int d);
int vlpclip_grad_both(const void* x_f16, int ldx, const void* y_f16, int ldy, const float* x_max,
                      const float* x_lg2l, const float* x_q, const float* y_max, const float* y_lg2l,
                      const float* y_q, int n_rows, int n_cols, int d, const float* scale,
                      int diag_shift, int n_global, float w_row, float w_col, const float* out_mul,
                      int dx_bf16, void* dx, int dy_bf16, void* dy, void* const* dy_owner_rows,
                      int n_owners, int rows_per_owner, float* dscale, void* workspace,
                      size_t workspace_bytes, void* stream);
/* ---- duplicate-caption-aware (masked) InfoNCE: SURVEY section 8 (f3) ----
 * row_ids [n_rows] / col_ids [n_cols]: int32 caption ids >= 0 (device).  A logit whose row and column
 * carry the same id is NOT a negative: unless it is the positive pair itself it is excluded from both
 * soft-max denominators (forward) and gets G = 0 (backward).  The mask is the one of the reference's
 * _get_mask (VisionLanguageModule.py:506-530: 0 where captions agree off the diagonal); the
 * reference's own use of it is deprecated code (:535-547), so "excluded" is this library's definition.
 * The forward entry is vlpclip_lse_fwd_fused(_f16) plus the ids (operand_f16 selects the operand type);
 * the backward entry is vlpclip_grad_both plus the ids, fed with the masked statistics. */
int vlpclip_lse_fwd_fused_masked(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                                 int d, int operand_f16, const float* scale, int diag_shift,
                                 const int* row_ids, const int* col_ids, float* row_max, float* row_l,
                                 float* diag, float* col_max, float* col_l, void* workspace,
                                 size_t workspace_bytes, void* stream);
int vlpclip_grad_both_masked(const void* x_f16, int ldx, const void* y_f16, int ldy, const float* x_max,
                             const float* x_lg2l, const float* x_q, const float* y_max,
                             const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                             const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                             const float* out_mul, int dx_bf16, void* dx, int dy_bf16, void* dy,
                             void* const* dy_owner_rows, int n_owners, int rows_per_owner, float* dscale,
                             const int* row_ids, const int* col_ids, void* workspace,
                             size_t workspace_bytes, void* stream);

/* host-only: the schedule of vlpclip_grad_both for np producer slots and nq >= np dT consumers
 * (np <= 0: the split used on the current device).  info[8] = {phases, nominal steps, dI partial
 * slots, runs, waves, np, nq, 0}; prod rows = {slot, step, row block, column tile, dI partial slot
 * or -1}; cons rows = {consumer, slot, step, row block, column tile, rank of the piece in its
 * column, pieces of the column} in the order each consumer works.  Used by the CPU test-suite. */
int vlpclip_grad_both_plan(int n_row_blocks, int n_col_tiles, int np, int nq, int* info, int* prod,
                           int max_prod, int* n_prod, int* cons, int max_cons, int* n_cons);

/* out[slot_elems] (fp32 or bf16) = out_mul * sum_s slots[s * slot_elems + .], s ascending */
int vlpclip_slot_sum(const float* slots, int n_slots, size_t slot_elems, const float* out_mul,
                     int out_bf16, void* out, void* stream);
/* peer windows: cudaMalloc + CUDA IPC export / import (handle64 = 64-byte cudaIpcMemHandle_t) */
int vlpclip_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int vlpclip_peer_open(const unsigned char* handle64, void** dev_ptr);
int vlpclip_peer_close(void* dev_ptr);
int vlpclip_peer_free(void* dev_ptr);

/* ---- retrieval metrics without the M x M similarity matrix (VisionLanguageModule.py:364-439) ----
 * Q [n_rows, d], K [n_cols, d]: L2-normalised embeddings as bf16 (row strides ldq / ldk elements).
 * Both entry points run the forward sweep's tile mainloop (tcgen05, S tile in TMEM) with a ranking
 * epilogue; similarities are never written to memory.  Order of two columns of a row: larger
 * similarity first, equal similarities by ascending column index (a stable descending sort).
 *   vlpclip_retrieval_ranks: rank[i] = number of columns ranked before column i in row i (needs
 *     n_cols >= n_rows).  recall@k of recall_at_k_on_image_text_retreival (:402-439) is
 *     mean(rank < k) for every k at once.
 *   vlpclip_retrieval_topk: idx[i, 0..k) (and optionally val) = the k best columns of row i,
 *     1 <= k <= 16; precision_at_k_on_image_embeddings (:364-400) takes k+1 of them, drops the
 *     first and compares labels.
 * workspace: vlpclip_retrieval_workspace_bytes(n_rows, n_cols, d, k) (k = 0 for the ranks). */
size_t vlpclip_retrieval_workspace_bytes(int n_rows, int n_cols, int d, int k);
int vlpclip_retrieval_ranks(const void* q_bf16, int ldq, const void* k_bf16, int ldk, int n_rows,
                            int n_cols, int d, int* rank, void* workspace, size_t workspace_bytes,
                            void* stream);
int vlpclip_retrieval_topk(const void* q_bf16, int ldq, const void* k_bf16, int ldk, int n_rows,
                           int n_cols, int d, int k, int* idx, float* val, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- prologue: emb = normalize(feat @ W) (VisionLanguageModule.py:448-453) ----
 * feat [n, f] fp32, W [f, d] fp32 (x @ W convention, not nn.Linear).
 * emb_f32 [n, d] fp32 (returned to the caller), emb_bf16 / emb_f16 [n, d] (operands of the
 * fused loss), inv_norm [n] = 1 / max(||u||, 1e-12).
 */
size_t vlpclip_project_workspace_bytes(int n, int f, int d);
int vlpclip_project_normalize_fwd(const float* feat, const float* w, int n, int f, int d,
                                  float* emb_f32, void* emb_bf16, void* emb_f16, float* inv_norm,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* du = (dE - E * rowsum(E * dE)) * inv_norm   (backward of F.normalize, :452-453) */
int vlpclip_normalize_bwd(const float* emb_f32, const float* d_emb, const float* inv_norm, int n,
                          int d, float* du, void* stream);

/* generic small GEMMs of the projection and its backward (tf32 tensor cores, fp32 accumulate):
 *   C[m, n] = A[m, k] @ B[k, n]      (trans_a = 0)
 *   C[m, n] = A[k, m]^T @ B[k, n]    (trans_a = 1)   -- dW = feat^T du
 *   trans_b = 1 reads B as [n, k]                    -- dfeat = du W^T
 * All operands row-major fp32 and read in place (no transposed copies); the contiguous extent of
 * each operand must be a multiple of 4 floats.  The workspace only holds split-K partial sums
 * (summed in a fixed order: results are bit-reproducible).
 */
int vlpclip_gemm_tf32(const float* a, const float* b, float* c, int m, int n, int k, int trans_a,
                      int trans_b, void* workspace, size_t workspace_bytes, void* stream);
size_t vlpclip_gemm_workspace_bytes(int m, int n, int k);

#ifdef __cplusplus
}
#endif
#endif /* VLPCLIP_H_ */
