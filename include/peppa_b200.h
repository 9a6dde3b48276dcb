/* peppa_b200 -- C ABI of the B200-native contrastive-scoring hot path.
 *
 * The reference (gchrupala/peppa) has no FFI layer: its seam is the Python module API of
 * pig/loss.py, pig/metrics.py, pig/triplet.py and pig/util.py:9-13.  The Python modules in
 * peppa_b200/ keep those signatures and call ONLY the entry points declared here (ctypes, see
 * INTEGRATION.md).  Plain pointers and sizes; no torch types.  All pointers are DEVICE
 * pointers unless stated otherwise; matrices are row-major with an element leading
 * dimension `ld*`.  Embedding rows may be bf16, fp16 or fp32 (`dtype`, a PB2_* code: the reference's
 * callers hand over fp32 tensors, fp16 under Lightning's `precision: 16` -- hparams_base.yaml:45,
 * pig/evaluation.py:70 -- or bf16): the row-wise kernels read the true values; the tensor-core kernels
 * (pb2_sim_*) take bf16 or fp16 operands natively and fp32 rows as their split-fp16 pair (pb2_split_f16:
 * contraction length 3 dim).  Operands are 16-byte aligned with dim % 64 == 0 (the Python side zero-pads
 * otherwise; zero padding changes neither dot products nor norms).  Every call only ENQUEUES work on `stream` (a cudaStream_t) and
 * returns a status: 0 = ok, non-zero = error, message via pb2_last_error().  Re-entrant; no
 * global state besides a cached driver entry point and per-device SM counts.
 */
#ifndef PEPPA_B200_H
#define PEPPA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_VERSION 2

#define PB2_OK 0
#define PB2_ERR_ARG 1
#define PB2_ERR_CUDA 2
#define PB2_ERR_UNSUPPORTED 3

/* element types accepted by the HBM-bound kernels */
#define PB2_BF16 0
#define PB2_F16 1
#define PB2_F32 2
/* one-byte gradient matrix (hinge indicator sums, exactly {0, 1, 2}) and its embedding operand: two 8-bit planes
 * per row, [hi (s8) | lo (u8)], of q = round(xhat * 32512) = 256 hi + lo (pb2_rows_quant_i8) */
#define PB2_U8 3
#define PB2_I8_PLANES 4

const char* pb2_last_error(void);
int pb2_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long pb2_launch_count(void);

/* ---- (c) triplet scoring: replaces pig/metrics.py:45-52 triplet_accuracy and the gathers of
 * pig/triplet.py:71-73,89-91.  out[t] = cos(a_t,p_t) - cos(a_t,n_t) (discrete == 0) or
 * (sign(.)+1)/2 in {0, 0.5, 1} (discrete != 0), F.cosine_similarity semantics (norms clamped
 * at 1e-8).  *_idx are optional int64 row indices (NULL = row t). */
int pb2_triplet_score(const void* anchor, const void* positive, const void* negative, const int64_t* anchor_idx,
                      const int64_t* positive_idx, const int64_t* negative_idx, int64_t n_triplets, int dim,
                      int64_t ld, int dtype, int discrete, float* out, void* stream);

/* ---- row statistics: replaces U.norm(2, dim=1, keepdim=True) of pig/util.py:11-12.
 * rinv[i] = 1/||x_i||_2 (no epsilon: a zero row gives +inf and NaN scores, like the reference),
 * norm[i] = ||x_i||_2; either output may be NULL.  x is bf16 / fp16 / fp32 (dtype). */
int pb2_row_norms(const void* x, int dtype, int64_t n, int dim, int64_t ld, float* rinv, float* norm, void* stream);

/* Split-fp16 tensor-core operand of an fp32 matrix x [n, dim]: per row, x' = x * rinv[i] * 2^e_i (rinv NULL = 1;
 * e_i puts the row's largest component in [512, 1024)), hi = fp16(x'), lo = fp16(x' - hi) (22 significant bits);
 * out [n, 3 dim] fp16 = [hi | lo | hi] (side 0, the X operand) or [hi | hi | lo] (side 1, the Y operand), and
 * scale_out[i] = 2^-e_i.  ONE pb2_sim_* call with dtype = PB2_F16, dim' = 3 dim and rinv_x / rinv_y = the two
 * scale_out vectors then accumulates <hi_x,hi_y> + <lo_x,hi_y> + <hi_x,lo_y> in fp32 and returns
 * <x_i, y_j> * rinv_x[i] * rinv_y[j] to ~2^-22 relative per product -- the matmul of pig/util.py:13 on fp32
 * inputs (which the reference normalises BEFORE its matmul, pig/util.py:11-12) without rounding them to bf16. */
int pb2_split_f16(const float* x, const float* rinv, int64_t n, int dim, int64_t ld, int side, void* out, int64_t ld_out,
                  float* scale_out, void* stream);

/* out[k] = <x[ix[k]], y[iy[k]]> * sx * sy with sx = rinv_x[ix[k]] (1 if rinv_x == NULL), same
 * for sy; ix / iy NULL = k.  The diagonal M_ii of pig/loss.py:43 and the positive's score of
 * pig/metrics.py:8-20.  dist_out (optional) = fl32(1 - out[k]); thr_out (optional) = the rank
 * threshold t_k: for every float s,  s >= t_k  <=>  fl32(1 - s) < dist_out[k]  (what pb2_sim_rank takes).
 * CUDA-core fp32 arithmetic: use it for the hinge diagonal; for ranking prefer pb2_sim_diag. */
int pb2_pair_dot(const void* x, const void* y, int dtype, const int64_t* ix, const int64_t* iy, const float* rinv_x,
                 const float* rinv_y, int64_t n, int dim, int64_t ldx, int64_t ldy, float* out, float* dist_out,
                 float* thr_out, void* stream);

/* Paired scores s_k = <x_k, y_k> * rinv_x[k] * rinv_y[k] through the SAME tcgen05 pipeline as the full
 * passes (only the diagonal tiles are visited), so that a gallery row duplicating the positive scores
 * bit-identically to it -- as in the reference, where positive and candidates come out of one GEMM
 * (pig/metrics.py:8) -- and "strictly closer" stays strict.  Outputs as pb2_pair_dot (each optional). */
int pb2_sim_diag(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t n, int dim, int dtype,
                 int64_t ldx, int64_t ldy, float* out, float* dist_out, float* thr_out, void* stream);

/* ---- (a)/(b) similarity kernels: S = X * Y^T on the tcgen05 tensor cores (X and Y both bf16 or both fp16
 * -- `dtype` = PB2_BF16 / PB2_F16, kind::f16 takes either natively; fp32 rows via pb2_split_f16 -- fp32
 * accumulate in TMEM), epilogue fused per entry point; S itself reaches HBM only in
 * pb2_sim_matrix.  X is [rows, dim], Y is [cols, dim]; s_ij = <x_i,y_j> * rinv_x[i] * rinv_y[j]
 * * scale (rinv_* == NULL means 1). */

/* pig/util.py:9-13 cosine_matrix (and the raw V A^T of pig/loss.py:19): out is fp32 [rows, cols]. */
int pb2_sim_matrix(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                   int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float* out, int64_t ld_out,
                   void* stream);

/* pig/metrics.py:7-40 with one target per query row: rank[i] += #{ j != pos_col[i] : s_ij >= pos_thr[i] }
 * = #{ j : fl32(1 - s_ij) < fl32(1 - s_pos) } with pos_thr from pb2_sim_diag / pb2_pair_dot.  rank (int32)
 * must be zeroed by the caller; column indices are offset by col_offset (sharded galleries).
 * No sort, no N x N store. */
int pb2_sim_rank(const void* q, const void* g, const float* rinv_q, const float* rinv_g, const float* pos_thr,
                 const int64_t* pos_col, int64_t rows, int64_t cols, int64_t col_offset, int dim, int dtype, int64_t ldq,
                 int64_t ldg, int32_t* rank, void* stream);

/* pig/metrics.py:54-77 resampled recall from ONE score matrix: scores = pb2_sim_matrix(references,
 * candidates) [G, G] fp32; idx = the n_samples x size subset draws (row-major int64, device);
 * rank[s, j] = #{ c != j : fl32(1 - scores[ix_j, ix_c]) < fl32(1 - scores[ix_j, ix_j]) } (int32 [n_samples, size]).
 * Replaces 500 small GEMMs + 50 000 argsort rows per evaluation call. */
int pb2_subset_rank(const float* scores, int64_t ld, const int64_t* idx, int n_samples, int size, int32_t* rank,
                    void* stream);

/* pig/loss.py:28-48 TripletLoss / contrastive, forward pass fused with the gradient matrix.
 * For local rows i (global row id row_offset + i) and columns j (global id col_offset + j), i != j globally:
 *   zc = margin + s_ij - diag_col[j],  zr = margin + s_ij - diag_row[i]
 *   row_cnt[i] += [zr >= 0]      col_cnt[j] += [zc >= 0]   (int32, caller-zeroed)
 *   loss_partial[cta] += ([zc >= 0] + [zr >= 0]) * s_ij    (fp32, one slot per CTA, deterministic);
 *     relu(zc) + relu(zr) summed = these partials + sum_j (margin - diag_col[j]) col_cnt[j]
 *     + sum_i (margin - diag_row[i]) row_cnt[i], completed by pb2_hinge_loss_terms
 *   gmat[i,j] = [zc >= 0] + [zr >= 0] in {0, 1, 2}                         (0 on the diagonal)
 *     as fp16 (g_dtype = PB2_F16, ld_g in elements) or as ONE BYTE per entry (g_dtype = PB2_U8: half the HBM
 *     traffic of the gradient matrix; consumed by the kind::i8 path of pb2_grad_gemm*)
 * gmat may be NULL (forward only).  |n_partials| = capacity of loss_partial (>= pb2_sim_grid()); the buffer
 * is cleared first unless n_partials is negative (the caller already did, see pb2_hinge_prep).
 * If rank != NULL the same pass also does pb2_sim_rank with the diagonal as the positive:
 * rank[i] += #{ j != i : s_ij >= pos_thr[i] } (loss and recall@k share one S pass). */
int pb2_sim_hinge(const void* x, const void* y, const float* rinv_x, const float* rinv_y, const float* diag_row,
                  const float* diag_col, int64_t rows, int64_t cols, int64_t row_offset, int64_t col_offset, int dim, int dtype,
                  int64_t ldx, int64_t ldy, float margin, float* loss_partial, int n_partials, int32_t* row_cnt,
                  int32_t* col_cnt, void* gmat, int g_dtype, int64_t ld_g, const float* pos_thr, int32_t* rank,
                  void* stream);

/* pig/loss.py:13-26 MILNCELoss: row-wise online log-sum-exp of s over all columns.
 * part_max / part_sum are [n_col_tiles * 2, rows] fp32 partials in the log2 domain
 * (pb2_sim_lse_parts() gives the first dimension); pb2_lse_merge folds them into lse[rows]
 * (natural log).  Column LSE = the same call with X and Y swapped. */
int pb2_sim_lse_parts(int64_t cols);
int pb2_sim_lse_rows(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                     int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float* part_max, float* part_sum,
                     void* stream);
int pb2_lse_merge(const float* part_max, const float* part_sum, int n_parts, int64_t rows, float* lse,
                  int accumulate, void* stream);

/* pig/loss.py:13-26 MILNCELoss, both directions of the log-sum-exp from ONE pass over the logits
 * (x = V A^T is both `x` and `x.permute(1,0,2)` of pig/loss.py:23): for logits with a known bound,
 * |scale * <x_i, y_j>| <= bound and bound * log2(e) <= 60, accumulate e_ij = 2^(logit_ij * log2(e) - M),
 * M = bound * log2(e), into row sums (row_part_sum: [pb2_sim_lse_parts(cols), rows]) and column sums
 * (col_part_sum: [pb2_sim_lse_col_parts(rows), cols], one partial per 32 rows, fixed order, no atomics).
 * pb2_lse_merge_const folds partials [n_parts, n] into lse[n] = (M + log2 sum_k part[k, i]) * ln 2
 * (log-added into lse when accumulate != 0).  Unbounded logits: pb2_sim_lse_rows twice. */
int pb2_sim_lse_col_parts(int64_t rows);
int pb2_sim_lse_both(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                     int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float bound,
                     float* row_part_sum, float* col_part_sum, void* stream);
/* The same pass with the ranking of pb2_sim_rank fused in -- loss statistics and recall@k from ONE pass over the
 * scores (pig/loss.py:13-26 + pig/metrics.py:23-40): rank[i] += #{ j != row_offset + i - col_offset :
 * fl32(fl32(<x_i, y_j> * rank_rinv_x[i]) * rank_rinv_y[j]) >= pos_thr[i] }, pos_thr from pb2_sim_diag.  The logits keep
 * rinv_x / rinv_y / scale (MILNCELoss does not re-normalise), the ranking takes its own 1 / ||row|| vectors
 * (pig/util.py:11-12; null = 1).  Counts are bit-identical to pb2_sim_rank's. */
int pb2_sim_lse_both_rank(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                          int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float bound,
                          float* row_part_sum, float* col_part_sum, const float* rank_rinv_x, const float* rank_rinv_y,
                          const float* pos_thr, int64_t row_offset, int64_t col_offset, int32_t* rank, void* stream);
int pb2_lse_merge_const(const float* part_sum, int n_parts, int64_t n, float bound, float* lse, int accumulate,
                        void* stream);

/* out[i] = log sum_k exp(parts[k * n + i]) (natural log): the cross-rank merge of column log-sum-exp
 * partials of a row-sharded gallery (each rank holds the LSE over its own rows). */
int pb2_lse_combine(const float* parts, int n_parts, int64_t n, float* out, void* stream);

/* MIL-NCE gradient matrix: gmat[i,j] = fp16( (exp(s_ij - den_row[i]) + exp(s_ij - den_col[j])) * 2^13 ). */
int pb2_sim_lse_grad(const void* x, const void* y, const float* rinv_x, const float* rinv_y, const float* den_row,
                     const float* den_col, int64_t rows, int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy,
                     float scale, void* gmat, int64_t ld_g, void* stream);

/* ---- backward GEMMs on the tensor cores: out[M, dim] (=|+=) alpha * op(G) * Z with G [g_rows, g_cols]
 * and Z 16-bit (PB2_F16 or PB2_BF16 each), fp32 accumulate, fp32 out -- or G one byte per entry (PB2_U8, values
 * {0, 1, 2}) with Z = PB2_I8_PLANES (pb2_rows_quant_i8; ld_g / ldz in bytes, dim % 256 == 0, contraction length
 * <= 32768 per call): tcgen05.mma.kind::i8 with exact s32 accumulation of the two planes, joined in fp32.  transpose == 0: out = G * Z
 * (M = g_rows, Z is [g_cols, dim]); transpose != 0: out = G^T * Z (M = g_cols, Z is [g_rows, dim]).
 * The autograd backward of torch.matmul in pig/util.py:13 / pig/loss.py:19. */
int pb2_grad_gemm(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, int transpose,
                  const void* z, int z_dtype, int dim, int64_t ldz, float alpha, int accumulate, float* out,
                  int64_t ld_out, void* stream);

/* The same product with a caller workspace (device, 256-byte aligned, pb2_grad_gemm_workspace() bytes, ZEROED
 * once after allocation; the kernel leaves its flag area zero).  With it, shapes whose whole output tiles do
 * not fill the last wave of SMs run stream-K: the contraction of a row block may be cut between two CTAs
 * and is completed through the workspace in a fixed order (deterministic, no atomics on `out`).  One
 * workspace serves one stream at a time.  workspace == NULL behaves like pb2_grad_gemm. */
int64_t pb2_grad_gemm_workspace(void);
int pb2_grad_gemm_ws(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, int transpose,
                     const void* z, int z_dtype, int dim, int64_t ldz, float alpha, int accumulate, float* out,
                     int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Both backward products of one gradient-matrix block in ONE launch (a batch-1k training step is launch
 * bound): out0[g_rows, dim] = alpha * G * Z0 (Z0 is [g_cols, dim]) and out1[g_cols, dim] = alpha * G^T * Z1
 * (Z1 is [g_rows, dim]).  Used when every 128 x 64 tile of both products fits the machine at once; larger
 * shapes run as two pb2_grad_gemm launches. */
int pb2_grad_gemm_dual(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z0,
                       const void* z1, int z_dtype, int dim, int64_t ldz0, int64_t ldz1, float alpha, float* out0,
                       float* out1, int64_t ld_out0, int64_t ld_out1, void* stream);

/* out = fp16(x * rinv) (rinv == NULL: plain conversion to fp16; x bf16 / fp16 / fp32): the embedding operand
 * of pb2_grad_gemm.  tcgen05 kind::f16 cannot mix fp16 and bf16 operands, and fp16 holds every
 * normalised bf16 embedding value with 3 extra significand bits. */
int pb2_rows_scale_f16(const void* x, int dtype, const float* rinv, int64_t n, int dim, int64_t ld, void* out,
                       int64_t ld_out, void* stream);

/* The kind::i8 form of the same operand, for a one-byte (PB2_U8) gradient matrix: out [n, 2 dim] bytes =
 * [hi plane (s8) | lo plane (u8)] of q = round(x * rinv * 32512) = 256 hi + lo (16 bits, one scale per tensor);
 * pass it to pb2_grad_gemm* as z with z_dtype = PB2_I8_PLANES and ldz = ld_out (bytes). */
int pb2_rows_quant_i8(const void* x, int dtype, const float* rinv, int64_t n, int dim, int64_t ld, void* out,
                      int64_t ld_out, void* stream);

/* y0 = T(x0 * coef[0]), y1 = T(x1 * coef[0]) (x fp32, coef on the device; T = out_dtype PB2_BF16 / PB2_F16 /
 * PB2_F32; n_elems per array, whole 16-byte output vectors): the backward of pig/loss.py's scalar losses, whose
 * gradients are produced in the forward and kept in fp32 -- autograd's grad_output (e.g. a GradScaler's 65536
 * under pig's `precision: 16`, hparams_base.yaml:45) is applied BEFORE the rounding to the inputs' dtype. */
int pb2_scale_pair(const float* x0, const float* x1, int64_t n_elems, int out_dtype, const float* coef, void* y0,
                   void* y1, void* stream);

/* Hinge finish (SURVEY 8a'): g_i = p_i + gdiag_i * rinv_y[i] * y_i,  gdiag_i = -(row_cnt[i] + col_cnt[i]),
 * p = G * Yhat from pb2_grad_gemm; grad_x[i] = coef * rinv_x[i] * (g_i - xhat_i <g_i, xhat_i>),
 * xhat = x * rinv_x (the Jacobian of the row normalisation in pig/util.py:11-12).
 * coef_dev (device scalar, may be NULL = 1) * coef_host multiplies the result. */
int pb2_hinge_finish(const float* p, int64_t ld_p, const void* x, const void* y, int dtype, const float* rinv_x,
                     const float* rinv_y, const int32_t* row_cnt, const int32_t* col_cnt, int64_t rows, int dim,
                     int64_t ldx, int64_t ldy, float coef_host, const float* coef_dev, float* grad_x,
                     int64_t ld_grad, void* stream);

/* Fused small-batch training step (pig/models.py:262 at batch ~1k is launch bound): ONE launch before
 * the similarity pass -- row norms of V and A, the diagonal score (pig/loss.py:43), fp16 normalised
 * copies vh/ah ([n, dim], ld = dim) for pb2_grad_gemm, zeroed counts and partials -- and ONE launch
 * after the gradient GEMMs: dV (rows of p_v) and dA (rows of p_a) through the normalisation Jacobian and
 * the diagonal term, plus the scalar loss = coef * (sum partials + sum_i (margin - diag_i)(row_cnt_i +
 * col_cnt_i)) (NaN when a row norm is zero); d_v / d_a are [n, dim] in out_dtype (PB2_F32/BF16/F16).
 * Between them: pb2_sim_hinge with n_partials passed
 * NEGATIVE (= "already zeroed", no memset) and two pb2_grad_gemm.  v / a are bf16 / fp16 / fp32 rows (dtype);
 * for fp32 rows pb2_hinge_prep also writes their split-fp16 tensor-core operands v_split / a_split
 * ([n, 3 dim] fp16) and scales scale_v / scale_a ([n] fp32), see pb2_split_f16; all NULL otherwise. */
int pb2_hinge_prep(const void* v, const void* a, int dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float* rinv_v,
                   float* rinv_a, float* diag, void* vh, void* ah, int32_t* row_cnt, int32_t* col_cnt,
                   float* loss_partial, int n_partials, void* v_split, void* a_split, float* scale_v, float* scale_a,
                   const float* rinv_v_in, const float* rinv_a_in, void* stream);
int pb2_hinge_finish2(const float* p_v, const float* p_a, const void* v, const void* a, int dtype, int64_t n, int dim, int64_t ldv,
                      int64_t lda, const float* rinv_v, const float* rinv_a, const float* diag, const int32_t* row_cnt,
                      const int32_t* col_cnt, const float* loss_partial, int n_partials, float margin, float coef,
                      float* loss_out, void* d_v, void* d_a, int out_dtype, void* stream);

/* The five launches above behind one call (one FFI crossing per training step): workspace is a 256-byte
 * aligned device buffer of pb2_hinge_step_workspace(n, dim, in_dtype) bytes; n <= 32768 (one gradient-matrix
 * block); v / a are bf16 / fp16 / fp32 rows (in_dtype).  rinv_v_in / rinv_a_in (optional, NULL = compute): the
 * 1/||row|| vectors a producer already holds -- pb2_project_normalize emits exactly these for its bf16 rows
 * (SURVEY 8f row 3), so the loss does not re-derive the norms of embeddings the encoder tail just normalised. */
int64_t pb2_hinge_step_workspace(int64_t n, int dim, int in_dtype);
int pb2_hinge_step(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float margin,
                   void* workspace, int64_t workspace_bytes, float* loss_out, void* d_v, void* d_a, int out_dtype,
                   const float* rinv_v_in, const float* rinv_a_in, void* stream);

/* The same step as an autograd pair (pig/models.py:262: `loss = self.loss(V, A)` ... `loss.backward()`): FOUR launches
 * for forward + backward instead of the five of pb2_hinge_step + pb2_scale_pair, and no fp32 gradient round trip.
 * pb2_hinge_forward = pb2_hinge_prep -> pb2_sim_hinge -> both gradient products, with the scalar loss folded by a spare
 * CTA of the product grid (a launch of its own when the products do not fit one grid): loss_out is final when it
 * returns.  What the backward needs -- G A^ and G^T V^ (fp32), 1/||row||, the indicator counts -- stays in `state`, a
 * 256-byte aligned device buffer of pb2_hinge_state_bytes(n, dim) bytes owned by the caller (one per forward whose
 * backward is still to come); `workspace` (pb2_hinge_forward_workspace bytes) is scratch and may be reused at once.
 * pb2_hinge_backward = ONE launch: d_v / d_a ([n, dim], out_dtype) = the gradients of the mean hinge loss times
 * grad_out[0] (device fp32 scalar, autograd's grad_output; NULL = 1), multiplied in fp32 BEFORE the rounding to
 * out_dtype -- an AMP loss scale reaches an fp16 gradient of ~1e-7 before the rounding does.  v / a: the rows the
 * forward saw.  Same bits as pb2_hinge_step followed by pb2_scale_pair.
 * state == NULL (state_bytes ignored): the loss alone, for a forward no backward will follow (validation) -- prep,
 * the hinge pass without a gradient matrix, the fold; workspace then needs pb2_hinge_step_workspace bytes. */
int64_t pb2_hinge_forward_workspace(int64_t n, int dim, int in_dtype);
int64_t pb2_hinge_state_bytes(int64_t n, int dim);
int pb2_hinge_forward(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float margin,
                      void* workspace, int64_t workspace_bytes, void* state, int64_t state_bytes, float* loss_out,
                      const float* rinv_v_in, const float* rinv_a_in, void* stream);
int pb2_hinge_backward(const void* state, int64_t state_bytes, const void* v, const void* a, int in_dtype, int64_t n, int dim,
                       int64_t ldv, int64_t lda, const float* grad_out, void* d_v, void* d_a, int out_dtype, void* stream);

/* MIL-NCE finish: grad_x[i] = coef * (p_i * 2^-13 - y_i) (coef = grad_out / N). */
int pb2_milnce_finish(const float* p, int64_t ld_p, const void* y, int dtype, int64_t rows, int dim, int64_t ldy,
                      float coef_host, const float* coef_dev, float* grad_x, int64_t ld_grad, void* stream);

/* MIL-NCE finish with K candidates per clip (pig/loss.py:19-25, x viewed as [N, N, K]):
 *   grad_x[r] = coef * ( p_r * 2^-13 - sum_{k < group} w[r*group + k] * y[(r*group + k) / y_div] ),
 * w = softmax over the K paired logits of a clip.  Video side: group = K, y_div = 1, y = audio rows;
 * audio side: group = 1, y_div = K, y = video rows. */
int pb2_milnce_finish_k(const float* p, int64_t ld_p, const void* y, int dtype, const float* w, int64_t rows, int group, int y_div,
                        int dim, int64_t ldy, float coef_host, const float* coef_dev, float* grad_x, int64_t ld_grad,
                        void* stream);

/* out[0] (=|+=) alpha * ( sum_k partials[k] + sum_k (margin - diag[k]) * cnt[k] ); either group may be
 * NULL.  Completes the hinge loss from pb2_sim_hinge's partials and indicator counts (double
 * accumulation, fixed order: deterministic). */
int pb2_hinge_loss_terms(const float* partials, int n_partials, const float* diag, const int32_t* cnt, int64_t n,
                         float margin, float alpha, float* out, int accumulate, void* stream);

/* Deterministic fixed-order sum of n fp32 partials, scaled: out[0] = alpha * sum. */
int pb2_sum_partials(const float* partials, int n, float alpha, float* out, void* stream);

/* MIL-NCE loss from the merged statistics: out[0] = mean_i( logaddexp(lse_row[i], lse_col[i]) - diag[i] ),
 * den[i] = logaddexp(lse_row[i], lse_col[i]). */
int pb2_milnce_loss(const float* lse_row, const float* lse_col, const float* diag, int64_t n, float* den,
                    float* out, void* stream);

/* pig/loss.py:41-48 contrastive(M) on a materialised square fp32 matrix (HBM-bound): forward loss
 * partials (+ optional gradient matrix dM scaled by coef).  loss_partial must hold n_partials
 * floats followed by 2*n int32 of scratch for the indicator counts. */
int pb2_contrastive_matrix(const float* m, int64_t n, int64_t ld, float margin, float* loss_partial, int n_partials,
                           float* grad_m, int64_t ld_grad, float coef_host, const float* coef_dev, void* stream);

/* ---- encoder tail: nn.Linear(n_in, n_out) followed by F.normalize(p=2, dim=1), the last
 * two stages of both reference encoders (pig/models.py:96-109, :130-150), as one tcgen05 kernel.
 *   y = x W^T + bias  (x [rows, n_in] bf16, W [n_out, n_in] bf16 in nn.Linear layout, bias fp32 or NULL)
 *   out = bf16( y / max(||y||, eps) )   [rows, n_out], n_out % 64 == 0, n_out <= 512, n_in % 64 == 0
 *   rinv[r] = 1 / ||out_r||  (of the ROUNDED row: what pb2_sim_* take as rinv_x / rinv_y; optional)
 *   norm[r] = ||y_r||        (for the backward's normalisation Jacobian; optional)
 * y itself never reaches HBM: one CTA holds all n_out features of its 128 rows in TMEM. */
int pb2_project_normalize(const void* x, const void* w, const float* bias, int64_t rows, int n_in, int n_out,
                          int64_t ldx, int64_t ldw, float eps, void* out, int64_t ld_out, float* rinv, float* norm,
                          void* stream);

/* ---- multi-GPU: one process per GPU on one node; rank r owns rows [r N/P, (r+1) N/P).
 *
 * Peer memory over NVLink / NVSwitch.  pb2_ipc_export: CUDA IPC handle (PB2_IPC_HANDLE_BYTES bytes) of the
 * allocation `ptr` lies in, plus ptr's byte offset inside it; ship both to the other ranks by any host channel.
 * pb2_ipc_open (on the receiving rank, its own device current): maps the allocation with lazy peer access and
 * returns its base (for pb2_ipc_close) and the peer pointer base + offset, usable by this device's kernels.
 * pb2_peer_reduce: out[i] = sum_q src[q][i] (fp32, fixed order q = 0 .. n_src-1, n_src <= PB2_MAX_PEERS, n_elems
 * a multiple of 4): the dV reduce-scatter of the sharded backward as ONE kernel of ours over peer memory -- the
 * owner of a row block pulls every rank's partial rows and sums them; no collective kernel runs beside the
 * tensor-core grids during the step.  `src` is a HOST array of device pointers (local or peer).  The caller orders
 * the launch after a collective that every rank enters once its partials are complete (pb2_nccl_colstat_merge). */
#define PB2_IPC_HANDLE_BYTES 64
#define PB2_MAX_PEERS 16
int pb2_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out);
int pb2_ipc_open(const void* handle, int64_t offset, void** base_out, void** ptr_out);
int pb2_ipc_close(void* base);
int pb2_peer_reduce(const void* const* src, int n_src, int64_t n_elems, float* out, void* stream);

/* NCCL steps of the sharded gallery for hosts that bring their own ncclComm_t (`comm`; the library binds NCCL at
 * run time -- dlsym on the process, else dlopen("libnccl.so.2") -- pb2_nccl_available() says whether it found it):
 *   pb2_nccl_gallery_allgather   rank r's rows (n_local x row_bytes raw bytes) -> all rows, rank-major (8e.1:
 *                                embeddings, 1/||row||, diagonal scores)
 *   pb2_nccl_colstat_merge       one NCCL group summing the column counts (int32 [n_total]), the scalar loss and
 *                                the recall hit counts (fp32 [n_hits]) over the ranks (8e.3); NULL = skip that one
 *   pb2_nccl_dv_reduce_scatter   dV partials [world * n_local, dim] fp32 -> this rank's summed rows (8e.4), for
 *                                hosts without peer access (pb2_peer_reduce is the NVLink path) */
int pb2_nccl_available(void);
int pb2_nccl_gallery_allgather(void* comm, const void* local_rows, int64_t n_local, int64_t row_bytes, void* full_out,
                               void* stream);
int pb2_nccl_colstat_merge(void* comm, int32_t* col_cnt, int64_t n_total, float* loss, float* hits, int n_hits, void* stream);
int pb2_nccl_dv_reduce_scatter(void* comm, const float* partial_full, int64_t n_local, int dim, float* out_local, void* stream);

/* number of CTAs the persistent similarity kernels launch on the current device */
int pb2_sim_grid(void);

/* ---- host side of the duration-matched triplet sampler (pig/triplet.py:99-104 over pig/util.py shuffled / grouped /
 * pairs).  No device work: these replay, on a COPY of Python's MT19937 state, exactly the draws the reference makes on
 * the global `random` generator, so a seeded evaluation draws the same triplets and leaves the generator in the same
 * state, without 1.5 million interpreter-level calls per 500 samples.  mt_state: the 625 words of random.getstate()[1]
 * (624 state words + the index), updated in place -- the caller hands it back through random.setstate().
 * pb2_host_random_doubles: out[i] = the next n values of random.random().
 * pb2_host_sample_pairs: groups g = items[group_start[g] .. group_start[g + 1]) (clips of one duration, in the order of
 * grouped()); per sample and group: xs = shuffled(group), then for each of pairs(xs) target, distractor =
 * random.sample(pair, 2); pos / neg [n_samples * sum_g floor(len_g / 2)] int64 in the reference's order. */
int pb2_host_random_doubles(uint32_t* mt_state, int64_t n, double* out);
int pb2_host_sample_pairs(uint32_t* mt_state, const int64_t* items, const int64_t* group_start, int64_t n_groups,
                          int64_t n_samples, int64_t* pos, int64_t* neg);

#ifdef __cplusplus
}
#endif
#endif /* PEPPA_B200_H */
