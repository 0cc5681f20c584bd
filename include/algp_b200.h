/* algp_b200 -- C ABI of the B200-native GP inference / information-gain path.
 *
 * The reference (sumitsk/algp) is pure Python with no FFI of its own; its hot
 * path is the NumPy / gpytorch call surface of models.py, utils.py and
 * agent.py.  These entry points are what a binding for that surface needs;
 * each cites the reference code it replaces.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - All matrix pointers are DEVICE pointers (fp64 unless stated), row-major,
 *    16-byte aligned, leading dimensions even.  Hyper-parameters arrive on the
 *    HOST as the reference stores them: log-lengthscales, log-outputscale
 *    (run.py:36-37) -- `kind` 0 = RBF, 1 = Matern nu=1.5 (models.py:216-222).
 *  - Factor-side dimensions are padded to multiples of 128 (`*_pad`); the
 *    padding of a covariance matrix is the identity (algp_kbuild does it).
 *  - `stream` is a cudaStream_t; every call is asynchronous on it.
 *  - Return value: 0 ok, 1 invalid argument, 2 CUDA error (see
 *    algp_last_cuda_error), 3 not positive definite, 4 unsupported.  There is
 *    no CPU fallback.
 */
#ifndef ALGP_B200_H
#define ALGP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int algp_version(void);
const char* algp_strerror(int code);
const char* algp_last_cuda_error(void);

/* ---- K1: kernel matrix -------------------------------------------------- */
/* GPR.cov_mat (models.py:161-181): out = s^2 k(x1,x2) [+ diag(diag_add)] [+ diag_scalar I],
 * written over [n1_pad x n2_pad] (zeros outside [n1 x n2]; ones on the padded
 * diagonal when pad_identity).  x2 == NULL means x2 = x1.  out_dtype 0 = fp64,
 * 1 = fp32 (the reference's dtype, utils.py:19).  If dot_vec != NULL also
 * writes dot_partial[r][t] = sum over column tile t of out[r][c]*dot_vec[c]
 * (t < algp_kbuild_col_tiles) -- the mean of utils.py:301 without re-reading K. */
int algp_kbuild(const double* x1, int64_t n1, const double* x2, int64_t n2, int d,
                const double* log_ls_host, double log_os, int kind,
                const double* diag_add, double diag_scalar, int pad_identity,
                void* out, int64_t n1_pad, int64_t n2_pad, int64_t ld, int out_dtype,
                const double* dot_vec, double* dot_partial, void* stream);
int algp_kbuild_col_tiles(int64_t n2_pad, int out_dtype);
/* out[r] = bias + scale * sum_t partial[r][t] (+ addvec[r]) */
int algp_rowsum(const double* partial, int64_t rows, int nt, double scale, double bias,
                const double* addvec, double* out, void* stream);
/* M[row_of_col[k]][k] += v : the sigma_n^2 of cov_matrix (agent.py:90) under a (location, base) gather */
int algp_scatter_add(double* M, int64_t ld, const int32_t* row_of_col, int64_t ncols, double v, void* stream);

/* ---- K2: factor and solves ----------------------------------------------- */
/* In-place blocked Cholesky of the lower triangle, A = L L^T; replaces
 * np.linalg.inv(cov_aa) (utils.py:300) together with algp_trtri.  Also writes
 * inv(L_jj) into the diagonal 128-blocks of Linv.  *info_dev = 0, or the
 * 1-based column at which the matrix stopped being positive definite. */
int algp_potrf(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int* info_dev, void* stream);
/* Diagonal-block kernel of algp_potrf: columns eliminated per barrier (1, 2 or 4; default 2).  A tuning /
 * testing knob: all three give the same factor up to rounding. */
int algp_set_potf2_rank(int r);
/* Completes Linv = L^-1 from its diagonal blocks (recursive doubling).
 * work: algp_trtri_work_doubles(npad) doubles. */
int algp_trtri(const double* L, int64_t npad, int64_t ld, double* Linv, int64_t ldi, double* work,
               int zero_upper, void* stream);
int64_t algp_trtri_work_doubles(int64_t npad);
/* out = M v over the lower triangle (beta = L^-1 y) */
int algp_gemv_lower(const double* M, int64_t n, int64_t ld, const double* v, double* out, void* stream);
/* out = M^T v over the lower triangle (alpha = L^-T beta); work: algp_gemv_work_doubles(n) */
int algp_gemv_lower_t(const double* M, int64_t n, int64_t ld, const double* v, double* out, double* work, void* stream);
int64_t algp_gemv_work_doubles(int64_t n);
/* out2[0] = log det A = 2 sum log L_ii (replaces slogdet, utils.py:193); out2[1] = |v|^2 if v */
int algp_logdet_sumsq(const double* L, int64_t n, int64_t ld, const double* v, double* out2, void* stream);
/* V = Ks L^-T using the triangular structure of Linv (cov_xa inv(cov_aa) cov_xa^T of
 * utils.py:300-305 is V V^T).  V may be NULL; rn_partial[m][t] (t < npad/64), if not
 * NULL, receives the row sums of V^2 per 64-column tile: var = k** - sum_t rn_partial. */
int algp_trmm_rt(const double* Ks, int64_t mpad, int64_t ldk, const double* Linv, int64_t npad, int64_t ldi,
                 double* V, int64_t ldv, double* rn_partial, void* stream);
/* C = beta C + alpha A B^T (A [mpad x kpad], B [npad x kpad]).  lower_only: the lower triangle
 * (diagonal included) is updated, strictly-upper 128-blocks are untouched, elements above the
 * diagonal inside diagonal blocks are unspecified. */
int algp_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                 int64_t mpad, int64_t npad, int64_t kpad, double alpha, double beta, int lower_only, void* stream);

/* ---- K2': TF32 mode of the variance path (tcgen05 tensor cores, TMEM accumulators) ---- */
/* fp64 matrix -> fp32 hi / lo planes with hi exactly representable in TF32 (x ~= hi + lo) */
int algp_split_tf32(const double* src, int64_t rows, int64_t cols, int64_t ld, float* hi, float* lo,
                    int64_t ldo, void* stream);
/* Same row-norm partials as algp_trmm_rt (rn_partial[m][t], t < npad/128) computed as the split-TF32
 * product K_hi L_hi^T + K_hi L_lo^T + K_lo L_hi^T with fp32 accumulation: the 1e-4 tier of the
 * variance (utils.py:305-308).  The L planes must be ALLOCATED with their row count rounded up to a
 * multiple of 256 (rows past npad are never used in the result). */
int algp_trmm_rt_tf32(const float* Khi, const float* Klo, int64_t mpad, int64_t ldk, const float* Lhi,
                      const float* Llo, int64_t npad, int64_t ldl, double* rn_partial, void* stream);

/* ---- K2'': fp64-grade variance on the INT8 tensor cores (tcgen05 kind::i8, exact s32 accumulation) ---- */
/* fp64 matrix -> `nslices` (2..8) signed base-128 digit planes of every row scaled by a power of two:
 * src[r][c] = row_scale[r] * sum_p digit_p[r][c] * 2^(-6-7p) (+ a remainder below 2^(-7 nslices) row_scale[r]).
 * The digits are written in the tensor core's operand layout, tile by tile (rows * cols * nslices bytes):
 *   [row tile of tile_rows][32-column k chunk][plane][row group of 8][k half][8 rows][16 B]
 * tile_rows = 128 for the left operand (K), 64 for the right one (Linv); rows % tile_rows == 0, cols % 32 == 0.
 * mask (or NULL): algp_i8_mask_bytes(rows, cols, tile_rows) bytes, 8-byte aligned; receives one occupancy byte
 * per (row tile, k chunk), bit p set iff plane p of that tile chunk holds a non-zero digit (row stride
 * round_up(cols / 32, 8) bytes).  The GEMMs skip the products, loads and whole k chunks that the masks prove zero:
 * exact, and a large saving when the operands decay away from a diagonal (kernel matrices of spatially ordered
 * points). */
int algp_split_i8(const double* src, int64_t rows, int64_t cols, int64_t ld, int nslices, int tile_rows,
                  int8_t* planes, double* row_scale, uint8_t* mask, void* stream);
int64_t algp_i8_mask_bytes(int64_t rows, int64_t cols, int tile_rows);
/* Same row-norm partials as algp_trmm_rt (rn_partial[m][t], t < npad/64) from the digit tiles of K [mpad x npad]
 * (tile_rows 128) and Linv [npad x npad] (tile_rows 64): the digit products are exact integer GEMMs, groups of
 * equal weight are summed in fp64 (error ~ 2^(-7 nslices) of the row scales): the fp64 tier of the variance
 * (utils.py:305-308) at several times the DMMA rate.  Kmask / Lmask: occupancy masks from algp_split_i8 or NULL.
 * npad <= 32768. */
int algp_trmm_rt_i8(const int8_t* Kt, const double* Kscale, const uint8_t* Kmask, int64_t mpad, const int8_t* Lt,
                    const double* Lscale, const uint8_t* Lmask, int64_t npad, int nslices, double* rn_partial,
                    void* stream);
/* algp_trmm_rt with V stored ([mpad x npad], row stride ldv) and, if rn_partial is not NULL, the same row-norm
 * partials, from the digit tiles: the W^T = Sigma_{:,B} L^-T build of the posterior state. */
int algp_trmm_rt_store_i8(const int8_t* Kt, const double* Kscale, const uint8_t* Kmask, int64_t mpad, const int8_t* Lt,
                          const double* Lscale, const uint8_t* Lmask, int64_t npad, int nslices, double* V, int64_t ldv,
                          double* rn_partial, void* stream);
/* C [mpad x npad] = alpha A B^T + beta C from the digit tiles of A [mpad x kpad] (tile_rows 128) and
 * B [npad x kpad] (tile_rows 64), fp64-grade with nslices = 8; transposed != 0 stores C^T ([npad x mpad], ldc its
 * row stride); lower_only skips the tiles entirely above the diagonal.  mpad % 128 == 0, npad % 64 == 0,
 * kpad % 32 == 0, kpad <= 32768. */
int algp_gemm_nt_i8(const int8_t* At, const double* Ascale, int64_t mpad, const int8_t* Bt, const double* Bscale,
                    int64_t npad, int64_t kpad, int nslices, double alpha, double beta, double* C, int64_t ldc,
                    int transposed, int lower_only, void* stream);
/* algp_potrf + algp_trtri in one call with the O(n^3) work on the INT8 tensor cores: recursive 2 x 2 splitting
 * (L21 = A21 Linv11^T, A22 -= L21 L21^T, Linv21 = -Linv22 L21 Linv11) whose products are exact digit GEMMs;
 * blocks of `base` rows or fewer (base % 128 == 0) use the DMMA kernels.  A's lower triangle becomes L, Linv
 * the full lower-triangular inverse; *info_dev as algp_potrf.  work: algp_potrf_inv_i8_work_bytes() bytes. */
int algp_potrf_inv_i8(double* A, int64_t npad, int64_t ld, double* Linv, int64_t ldi, int nslices, int64_t base,
                      void* work, int64_t work_bytes, int* info_dev, void* stream);
int64_t algp_potrf_inv_i8_work_bytes(int64_t npad, int nslices, int64_t base);

/* ---- K4: marginal-likelihood gradient (GPR.fit, models.py:145-158) ---------- */
/* Ainv (lower triangle) = Linv^T Linv = A^-1 */
int algp_potri_lower(const double* Linv, int64_t npad, int64_t ldi, double* Ainv, int64_t lda, void* stream);
/* grad_out[d+2] = 0.5 tr((alpha alpha^T - A^-1) dA/dtheta), theta = (log lengthscale[d], log outputscale,
 * log noise); K and dK/dtheta are rebuilt on the fly.  The gradient of the reference's loss -mll/N is
 * -grad_out/N.  work: algp_mll_grad_work_doubles(n) doubles. */
int algp_mll_grad(const double* x, int64_t n, int d, const double* log_ls_host, double log_os, int kind,
                  double noise, const double* alpha, const double* Ainv, int64_t lda, double* work,
                  double* grad_out, void* stream);
int64_t algp_mll_grad_work_doubles(int64_t n);

/* ---- K3 / K5: information-gain scoring ------------------------------------ */
/* Entropy H(S1) of `B` candidate sets of `k` slots (k <= 128; idx -1 = empty
 * slot, duplicates idempotent) against the factored base set: replaces the
 * per-candidate slogdet loops of Agent.greedy / Agent.best_path
 * (agent.py:317-347, 373-400).  Wt [n x ldw] = Sigma_{:,B} L^-T (first `ncols`
 * columns valid, the rest zero up to a multiple of 16), X [n x d] the field
 * coordinates, pi0[n] the base precisions, delta[B x k] (or delta_scalar) the
 * precision each slot adds, skip[n] (or NULL) marks locations whose slots add
 * nothing (already mobile-sampled: the flag is boolean, agent.py:377).  scores[c] = H_base + n_new*CONST +
 * 0.5*(logdet(I + D P_CC D) - sum(log(pi0+delta) - [pi0>0] log pi0)). */
int algp_score_sets(const double* Wt, int64_t ldw, int64_t ncols, const double* X, int d,
                    const double* log_ls_host, double log_os, int kind, double noise,
                    const double* pi0, const int32_t* idx, const double* delta, double delta_scalar,
                    const uint8_t* skip, int k, int64_t B, double H_base, double* scores, void* stream);
/* The same scores for long paths: 128 < k <= 2048 slots per candidate (the reference scores paths of "tens to
 * hundreds" of mobile locations, agent.py:373-400).  The k x k conditional covariance of a candidate lives in
 * `work` instead of shared memory; k <= 128 ignores work.  algp_score_sets_large_work_doubles(k, B) is the PREFERRED
 * size (one matrix per resident CTA); any work_doubles >= kk (kk + 1), kk = k rounded up to 8, is accepted and runs
 * correspondingly fewer CTAs, so a caller can bound the scratch by the memory it has. */
int algp_score_sets_large(const double* Wt, int64_t ldw, int64_t ncols, const double* X, int d,
                          const double* log_ls_host, double log_os, int kind, double noise, const double* pi0,
                          const int32_t* idx, const double* delta, double delta_scalar, const uint8_t* skip, int k,
                          int64_t B, double H_base, double* scores, double* work, int64_t work_doubles, void* stream);
int64_t algp_score_sets_large_work_doubles(int k, int64_t B);
/* The same scores for k <= 8 with a workspace, which lets the library pick the launch structure:
 *  - small batches (up to ~1.5 x the resident warp slots, ~5000 sets): every candidate is scored by 2 or 4 independent warps,
 *    each over a slice of the columns; fragments and per-candidate arrival counters live in `work`, the warp that
 *    arrives last finishes the candidate (no barrier, no extra launch);
 *  - one call that streams >= ~2 GB of rows: the columns of Wt are swept in L2-sized chunks (random sets re-read every
 *    row of Wt many times; a chunk's slice of Wt, n_rows x chunk x 8 bytes, fits the L2 where whole rows do not) by
 *    ONE persistent launch whose SMs own their candidates for the whole call and keep the partial Grams in shared
 *    memory between chunks (1.22-1.29 ms on configs[2] against 1.51-1.67 for the single launch); the older form --
 *    one launch per chunk, accumulator fragments parked in `work` -- is kept behind the switches below (1.45-1.55 ms);
 *  - else the plain single launch.
 * work: algp_score_sets_tiled_work_doubles(B) doubles, 16-byte aligned, ZERO-FILLED when allocated and then left to
 * the library (its head holds arrival counters that are never reset); one workspace per stream.
 * algp_set_score_tile_cols(c): c > 0 forces a sweep in chunks of c columns (multiple of 64), -1 the plain single
 * launch, 0 the default policy.  algp_set_score_resident(m): how a chunked sweep runs -- 1 the persistent launch, -1
 * one launch per chunk, 0 (default) persistent when the library derived the chunk, per-chunk launches when
 * algp_set_score_tile_cols forced it.  Both are process-wide tuning / test switches. */
int algp_score_sets_tiled(const double* Wt, int64_t ldw, int64_t ncols, int64_t n_rows, const double* X, int d,
                          const double* log_ls_host, double log_os, int kind, double noise, const double* pi0,
                          const int32_t* idx, const double* delta, double delta_scalar, const uint8_t* skip, int k,
                          int64_t B, double H_base, double* scores, double* work, int64_t work_doubles, void* stream);
int64_t algp_score_sets_tiled_work_doubles(int64_t B);
/* Kernel launches one algp_score_sets_tiled call of this shape makes (one per column chunk in the per-chunk form, one
 * per ~68 000 candidates in the persistent form);
 * <= 0 for an invalid shape.  For callers that count launches. */
int algp_score_sets_tiled_launches(int k, int64_t B, int64_t ncols, int64_t n_rows);
int algp_set_score_tile_cols(int cols);
int algp_set_score_resident(int mode);
/* The same scores from a RESIDENT posterior covariance of the base set, P = Sigma + sigma_n^2 I - Wt Wt^T
 * (lower triangle of an [n x ldp] matrix; build it with algp_kbuild + algp_gemm_nt / algp_gemm_nt_i8, lower_only):
 * a candidate reads its k(k+1)/2 entries P[c_i][c_j] instead of k rows of Wt.  The Schur complement of
 * agent.py:373-400 / utils.py:188-194 is then a gather; pays off when one base set is scored many times.
 * k <= 128 (ALGP_ERR_UNSUPPORTED beyond: use algp_score_sets_large). */
int algp_score_sets_cov(const double* P, int64_t ldp, const double* pi0, const int32_t* idx, const double* delta,
                        double delta_scalar, const uint8_t* skip, int k, int64_t B, double H_base, double* scores,
                        void* stream);
/* Keep a resident P current across commits: P[i][j] -= sum_{c<k} Wt[i][col0+c] Wt[j][col0+c] on the lower triangle,
 * for the k (<= algp_cov_downdate_max_cols() = 32) columns Wt gained through algp_append / algp_append_block since P
 * was last current.  One pass over the lower triangle of P (HBM bound) instead of a rebuild (kernel matrix + SYRK). */
int algp_cov_downdate(double* P, int64_t ldp, int64_t n, const double* Wt, int64_t ldw, int64_t col0, int k, void* stream);
int algp_cov_downdate_max_cols(void);
/* greedy utilities for every location (k = 1 closed form, agent.py:341) */
int algp_greedy_utilities(const double* diagP, const double* pi, const uint8_t* is_static, double d_static,
                          int64_t n, double* ut, void* stream);
/* np.argmax semantics (first maximum, agent.py:349,402): out_pair = {double value; int64 index + idx_offset}.
 * work: algp_argmax_work_bytes() bytes. */
int algp_argmax(const double* x, int64_t n, int64_t idx_offset, void* out_pair, void* work, void* stream);
int64_t algp_argmax_work_bytes(void);
/* *bad_count (device int64) = number of candidate slots outside [-1, n) in idx[count] (-1 = empty slot); those
 * slots are overwritten with -1.  The reference indexes NumPy arrays with the path lists (agent.py:377) and raises
 * IndexError for such an entry; the scoring kernels do not range-check their slots, so the Python mirror runs this
 * pass over its device copy of a caller's slot array before scoring it and reads the count back together with the
 * winner (then raises IndexError). */
int algp_check_indices(int32_t* idx, int64_t count, int64_t n, int64_t* bad_count, void* stream);
/* ---- winner exchange between the GPUs of one box over NVLink peer memory (csrc/p2p.cu) ----
 * Replaces the NCCL all-gather of one 16-byte {score, global index} pair per rank that follows the sharded scoring
 * step: the last argmax kernel stores the rank's pair into a mailbox in every peer's memory, waits for
 * the peers' pairs in its own and reduces them with the first-maximum rule.
 * algp_p2p_create: cudaMalloc + zero a mailbox of `bytes` (algp_p2p_mailbox_bytes(world)) and export its 64-byte
 * cudaIpc handle; algp_p2p_open maps a peer's handle; close / destroy undo them. */
int64_t algp_p2p_mailbox_bytes(int world);
int algp_p2p_create(int64_t bytes, void** local_ptr, void* handle64);
int algp_p2p_open(const void* handle64, void** peer_ptr);
int algp_p2p_close(void* peer_ptr);
int algp_p2p_destroy(void* local_ptr);
/* np.argmax of this rank's block x[n] (global ids = position + idx_offset; n = 0 = empty shard) exchanged with the
 * `world` ranks whose mailboxes are listed in peers_dev (device array of `world` pointers, entry r = rank r's mailbox as
 * mapped here).  out3 (device, 24 bytes) = {double value; int64 index; int64 status (0 ok, 1 = a peer timed out)},
 * identical on every rank.  epoch = 1, 2, 3, ... in step on all ranks.  work: algp_argmax_work_bytes() bytes. */
int algp_argmax_exchange(const double* x, int64_t n, int64_t idx_offset, void* work, const void* peers_dev, int rank,
                         int world, int64_t epoch, double timeout_ms, void* out3, void* stream);
/* Commit an acquisition at location *j_dev with precision increment delta: appends column
 * `ncols` to Wt, downdates diagP, raises pi[j] (and is_static[j] if mark_static).
 * work: algp_append_work_doubles(n) doubles. */
int algp_append(double* Wt, int64_t ldw, int64_t ncols, const double* X, int64_t n, int d,
                const double* log_ls_host, double log_os, int kind, double noise,
                double* diagP, double* pi, uint8_t* is_static, const void* j_dev, double delta,
                int mark_static, double* work, void* stream);
int64_t algp_append_work_doubles(int64_t n);
/* Commit k (1..16) DISTINCT locations idx[k] (device int64) at once, e.g. the mobile readings of the chosen path
 * (agent.py:179-192): the same k new columns of Wt, diagP, pi and flags as k successive algp_append calls in that
 * order, but Wt is read once instead of k times.  delta[k] on the device, or NULL for delta_scalar.
 * work: algp_append_block_work_doubles() doubles.  Columns [ncols, ncols rounded up to 16) of Wt must be zero (they
 * are in a zero-filled Wt that only gains columns through these calls; the scoring kernels read the same tail).
 * With 32-byte aligned rows (ldw % 16 == 0) the pass over Wt runs on DMMA fragments; algp_set_append_block_scalar(1)
 * forces the scalar pass (tests / tuning, process-wide). */
int algp_append_block(double* Wt, int64_t ldw, int64_t ncols, const double* X, int64_t n, int d,
                      const double* log_ls_host, double log_os, int kind, double noise, double* diagP, double* pi,
                      uint8_t* is_static, const void* idx_dev, int k, const double* delta_dev, double delta_scalar,
                      int mark_static, double* work, void* stream);
int64_t algp_append_block_work_doubles(void);
int algp_set_append_block_scalar(int on);

/* ---- mutual-information criterion (agent.py:330-339, 388-397) ---------------------- */
/* out[c] = sum_{r>=c} M[r][c]^2 = diag(A^-1) from the inverse factor; work: algp_colsumsq_work_doubles(n) */
int algp_colsumsq_lower(const double* M, int64_t n, int64_t ld, double* out, double* work, void* stream);
int64_t algp_colsumsq_work_doubles(int64_t n);
/* Rank-1 maintenance of diag(A^-1) and logdet A across greedy picks (agent.py:330-339 recomputes two n x n slogdets
 * per candidate): col = column j of the ORIGINAL inverse, U [>= t+1 rows x ldu] / coef[t+1] the earlier corrections
 * (current inverse = A0^-1 - sum_s coef[s] U[s] U[s]^T).  mode 0: row / column j is deleted from A (logdet += log c_j);
 * mode 1: A += delta e_j e_j^T (logdet += log1p(delta c_j)); c = column j of the current inverse.  diag[n] and the device
 * scalar logdet are updated in place, U[t] = c and coef[t] are appended.  t < 64. */
int algp_inv_rank1_update(const double* col, int64_t n, double* U, int64_t ldu, double* coef, int t, int64_t j, int mode,
                          double delta, double* diag, double* logdet, void* stream);
/* Per candidate c: out3[c] = {logdet [A2^-1]_CC over its brand-new locations, their count,
 * sum log|Delta| + logdet|Delta^-1 + [A3^-1]_CC| over all its active locations}.  inv2 / inv3 are
 * the inverses (lower triangle valid, algp_potri_lower) of Sigma_AbarAbar and Sigma + D; pos2[n] maps a
 * location to its row in inv2 (-1 = sampled); skip[n] marks already-mobile locations (no-ops). */
int algp_mi_terms(const double* inv2, int64_t ld2, const int32_t* pos2, const double* inv3, int64_t ld3,
                  const int32_t* idx, int k, int64_t B, const uint8_t* skip, double delta_new, double delta_old,
                  double* out3, void* stream);
/* algp_mi_terms for 128 < k <= 2048 (scratch: algp_mi_terms_large_work_doubles(k, B) doubles preferred, any
 * work_doubles >= k (k + 1) accepted: fewer CTAs run) */
int algp_mi_terms_large(const double* inv2, int64_t ld2, const int32_t* pos2, const double* inv3, int64_t ld3,
                        const int32_t* idx, int k, int64_t B, const uint8_t* skip, double delta_new, double delta_old,
                        double* out3, double* work, int64_t work_doubles, void* stream);
int64_t algp_mi_terms_large_work_doubles(int k, int64_t B);

/* ---- growing-prefix posteriors (agent.py:497-518) ------------------------------------- */
/* out[i][m] = {sum_j V[m][j] beta[j], sum_j V[m][j] gamma[j], sum_j V[m][j]^2} over j < prefix[i]
 * (prefix ascending, device int32): mean and variance of the GP fitted to the first prefix[i] training
 * rows, for every i, from ONE factorisation of the full ordered set. */
int algp_prefix_reduce(const double* V, int64_t ldv, int64_t rows, const double* beta, const double* gamma,
                       const int32_t* prefix_dev, int nprefix, double* out, void* stream);

/* ---- diagnostics (bench.py; not on the product path) ----
 * L2 -> SM delivery rate: `bytes` of buf (L2-sized, 16-byte aligned) read `passes` times by sms x ctas_per_sm CTAs,
 * every CTA a different part, with the 16-byte no-allocate loads of the scoring kernels.  Time it with CUDA events:
 * bytes x passes / time is the roof of the row-streaming scoring kernel. */
int algp_probe_l2_read(const void* buf, int64_t bytes, int passes, int ctas_per_sm, void* sink, void* stream);

/* ---- path enumeration feeding Agent.best_path (host code, no GPU work) ---------------------------
 * The expansion-tree search of FieldEnv.get_all_paths (env.py:197-310) on a planning graph given in CSR form:
 * node_rc[n][2] = (row, col) of every node, neighbours adj[adj_ptr[v] .. adj_ptr[v+1]) in the graph's own
 * adjacency order, eidx[eidx_ptr[e] .. eidx_ptr[e+1]) = field locations sampled along directed entry e (the
 * edge's `indices`, the same list in both directions).  Breadth-first expansion from start_node with the given
 * heading, U-turns forbidden, children pruned by g + bounding-box lower bound > least_cost + slack, equal
 * (pose, heading, visited, g) children merged; every root path of the expansion DAG to a node that visited all
 * waypoints is returned in networkx.all_shortest_paths order.  n_waypoints <= 64.  *handle_out owns the result
 * until algp_paths_free. */
int algp_paths_enumerate(int32_t n_nodes, const int32_t* node_rc, const int64_t* adj_ptr, const int32_t* adj,
                         const int64_t* eidx_ptr, const int32_t* eidx, int32_t start_node, int32_t heading_r,
                         int32_t heading_c, const int32_t* waypoint_nodes, int32_t n_waypoints, double least_cost,
                         double slack, int64_t max_tree_nodes, void** handle_out);
/* sizes[0..5] = {paths, total path nodes, total indices, tree nodes, merged children, longest index list} */
int algp_paths_sizes(const void* handle, int64_t* sizes, double* least_cost);
/* ragged copies: path_ptr[paths+1] / path_nodes, idx_ptr[paths+1] / idx, cost[paths]; any pointer may be NULL */
int algp_paths_fetch(const void* handle, int64_t* path_ptr, int32_t* path_nodes, int64_t* idx_ptr, int32_t* idx,
                     double* cost);
/* slots[paths][k] (-1 = empty): the candidate matrix algp_score_sets / Agent.best_path take */
int algp_paths_fill_slots(const void* handle, int32_t* slots, int64_t k);
int algp_paths_free(void* handle);

#ifdef __cplusplus
}
#endif
#endif
