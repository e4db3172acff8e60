/* fgp.h — C-ABI of libfgp_sm100.so: the B200 (sm_100a) replacement for friedrich's dense-algebra hot path.
 *
 * The reference (Rust crate friedrich v0.5.1) has no FFI; the seam this ABI replaces is the crate-private boundary
 * between `gaussian_process` and `algebra` + nalgebra.  Each entry point below names the reference code it stands in
 * for (paths relative to the reference root).  INTEGRATION.md shows the `extern "C"` block and the modified
 * `algebra` module a friedrich maintainer would add.
 *
 * Conventions
 *   - every matrix is f64, COLUMN-major with an explicit leading dimension (nalgebra DMatrix / EMatrix::as_matrix,
 *     src/algebra/extendable_matrix.rs:52-55); inputs are n x d with one sample per ROW (src/conversion/mod.rs);
 *   - the caller owns all host buffers; the handle owns all device memory; nothing is retained from host pointers;
 *   - `y_resid` is training_outputs MINUS prior(X) (src/gaussian_process/mod.rs:156): priors stay on the host;
 *   - every function returns an fgp_status; nothing aborts or throws across the boundary. The reference's panics
 *     (src/algebra/mod.rs:85,90; src/gaussian_process/mod.rs:203,263,345) are mapped by the host shim;
 *   - a handle is bound to one GPU and one host thread at a time (`&mut self` entry points: fit/refit/add_samples/
 *     set_outputs; the `&self` ones — predict*, likelihood, lml_gradient — serialise on an internal mutex);
 *   - there is NO CPU fallback: without a CUDA device fgp_create fails with FGP_ERR_CUDA.
 */
#ifndef FGP_H
#define FGP_H

#include <stddef.h>
#include <stdint.h>

#include "fgp_kernel_desc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgp_model fgp_model; /* opaque: device X (row-major, padded), y, alpha, L, inverse diagonal blocks */

typedef enum fgp_status {
    FGP_OK = 0,
    FGP_ERR_NOT_POSDEF = 1,   /* Cholesky pivot <= 0 / NaN and no valid substitute; see fgp_failed_column */
    FGP_ERR_BAD_ARG = 2,      /* dimension mismatch, negative noise (mod.rs:150), null pointer ... */
    FGP_ERR_BAD_KERNEL = 3,   /* malformed or unsupported kernel descriptor */
    FGP_ERR_CUDA = 4,         /* CUDA runtime error (message in fgp_last_error) */
    FGP_ERR_NOT_FITTED = 5,   /* predict/likelihood/... before fgp_fit */
    FGP_ERR_COMM = 6          /* multi-GPU transport error */
} fgp_status;

/* lifecycle --------------------------------------------------------------------------------------------------- */
int fgp_create(int device, fgp_model** out);
int fgp_destroy(fgp_model* m);
const char* fgp_last_error(const fgp_model* m);
const char* fgp_version(void);

/* fit ----------------------------------------------------------------------------------------------------------
 * fgp_fit            make_cholesky_cov_matrix (src/algebra/mod.rs:59-92) as called by GaussianProcess::new
 *                    (src/gaussian_process/mod.rs:158-159): uploads X (n x d, ld = ldx) and y_resid, assembles the
 *                    Gram lower triangle + noise^2 I, factors it. has_eps/eps = cholesky_epsilon (mod.rs:67-73).
 * fgp_set_inputs     EMatrix::new(training_inputs) only (mod.rs:147): makes X resident without fitting, so that the
 *                    builder's heuristic_fit (builder.rs:193-196 -> fgp_mean_pair_distance) can run before the fit.
 * fgp_refit          same as fgp_fit, on the X (and y) already resident (optimiser loops optimizer.rs:133-136, :267-270 and
 *                    fit_parameters mod.rs:426-429).
 * fgp_set_outputs    EVector::assign after a prior refit (mod.rs:420, extendable_matrix.rs:107-111).
 * fgp_add_samples    EMatrix/EVector::add_rows + add_rows_cholesky_cov_matrix (mod.rs:181-189, algebra/mod.rs:97-126):
 *                    k sequential insert_column(end) == block update of the factor; no failure check in the
 *                    reference (sqrt of a negative gives NaN) — here a non-positive pivot returns FGP_ERR_NOT_POSDEF,
 *                    or takes the substitute when has_eps is set (the reference's insert_column never consults
 *                    cholesky_epsilon; with has_eps = 0 the behaviour differs only where the reference produces NaN).
 *                    On failure the handle keeps the OLD sample count and must be refitted (fgp_refit) before use.
 */
int fgp_set_inputs(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d);
int fgp_fit(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
            const fgp_kernel_desc* kernel, double noise, int has_eps, double eps);
int fgp_refit(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int has_eps, double eps);
int fgp_set_outputs(fgp_model* m, const double* y_resid, int64_t n);
int fgp_add_samples(fgp_model* m, const double* Xnew, int64_t ldx, int64_t k, const double* ynew_resid,
                    const fgp_kernel_desc* kernel, double noise, int has_eps, double eps);
int64_t fgp_failed_column(const fgp_model* m); /* 0-based column of the failed pivot of the last fit, -1 if none */
int64_t fgp_num_samples(const fgp_model* m);
int64_t fgp_num_dims(const fgp_model* m);

/* predict ------------------------------------------------------------------------------------------------------
 * Xq is q x d (ld = ldq). Means are returned WITHOUT the prior (the shim adds prior(Xq), mod.rs:238-241).
 * fgp_predict_mean      GaussianProcess::predict                 (mod.rs:226-244)
 * fgp_predict_var       GaussianProcess::predict_variance        (mod.rs:248-273)
 * fgp_predict_mean_var  GaussianProcess::predict_mean_variance   (mod.rs:290-326)
 * fgp_predict_cov       predict_covariance (mode 0, mod.rs:329-350) / covariance of sample_at (mode 1, mod.rs:371-384);
 *                       cov is q x q column-major, ld = ldc; mean_wo_prior may be NULL.
 */
int fgp_predict_mean(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                     double* mean_wo_prior);
int fgp_predict_var(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q, double* var);
int fgp_predict_mean_var(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                         double* mean_wo_prior, double* var);
int fgp_predict_cov(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q, int mode,
                    double* cov, int64_t ldc, double* mean_wo_prior);

/* model selection ----------------------------------------------------------------------------------------------
 * fgp_likelihood     GaussianProcess::likelihood (mod.rs:196-220), formula as coded (penalty is sum ln|k(x,x)+noise^2|).
 * fgp_lml_gradient   scaled != 0: scaled_gradient_marginal_likelihood (optimizer.rs:159-203) -> *scale_out and P grads
 *                    scaled == 0: gradient_marginal_likelihood (optimizer.rs:24-60) -> P grads followed by the noise grad.
 *                    Replaces covmat_cholesky.inverse() + make_gradient_covariance_matrices (algebra/mod.rs:129-155);
 *                    the P gradient matrices are never materialised.
 * fgp_mean_pair_distance  fit_bandwidth_mean (src/parameters/kernel.rs:94-113) on the resident training inputs.
 */
int fgp_likelihood(fgp_model* m, const fgp_kernel_desc* kernel, double noise, double* out);
int fgp_lml_gradient(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int scaled, double* scale_out,
                     double* grads);
int fgp_mean_pair_distance(fgp_model* m, double* out);

/* state transfer (serde feature, mod.rs:58; tests) --------------------------------------------------------------
 * fgp_download_factor  L into a caller buffer, n x n column-major, strict upper triangle set to NaN exactly like
 *                      the matrix nalgebra's Cholesky keeps (algebra/mod.rs:67).
 * fgp_download_alpha   K^-1 y_resid (n values).
 */
int fgp_download_factor(fgp_model* m, double* L, int64_t ldl);
int fgp_download_alpha(fgp_model* m, double* alpha);
/* fgp_upload_state     the inverse of the two downloads — deserialising a saved GaussianProcess (serde feature: mod.rs:58,
 *                      EMatrix / EVector extendable_matrix.rs:14,62; nalgebra's Cholesky stores the n x n matrix whose lower
 *                      triangle is the factor): X (n x d, ld = ldx), y_resid and L (n x n, ld = ldl; the strict upper triangle
 *                      is ignored, NaN allowed) become resident WITHOUT refitting; the inverse diagonal blocks, alpha and
 *                      L^-1 y that the device path caches are rebuilt from L.  The handle then behaves as after fgp_fit.
 * fgp_inverse_columns  selected columns of K^-1 (`covmat_cholesky.inverse()`, optimizer.rs:32, :169) as left on the device
 *                      by the last fgp_lml_gradient call; out is n x ncols, ld = ldo.  Diagnostics / tests.
 * fgp_factor_digest    three deterministic sums over the lower triangle of the resident factor (sum, sum of squares,
 *                      position-weighted sum): bitwise-equal factors give bitwise-equal digests, so ranks of a sharded fit
 *                      (and a sharded against a single-GPU fit) can be compared without moving 8 n^2 bytes. out[3]. */
int fgp_factor_digest(fgp_model* m, double* out);
int fgp_upload_state(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
                     const double* L, int64_t ldl);
int fgp_inverse_columns(fgp_model* m, const int64_t* cols, int64_t ncols, double* out, int64_t ldo);

/* Cholesky of a caller-supplied SPD matrix with the same device factorisation: MultivariateNormal::new
 * (src/gaussian_process/multivariate_normal.rs:54-59, `covariance.cholesky().expect(..).unpack()`). A is n x n
 * column-major (lower triangle read) and is overwritten by L with the strict upper triangle zeroed. */
int fgp_cholesky_lower(int device, double* A, int64_t lda, int64_t n, int64_t* failed_col);

/* multi-GPU fit (no reference counterpart: the crate is single-threaded) ---------------------------
 * One process per GPU.  The 512-column panels of the covariance matrix are owned block-cyclically; each rank assembles
 * only its own panels (algebra/mod.rs:70-79 restricted to them), the owner of a panel factors it and ncclBroadcast()s it
 * over NVLink, every rank applies it to the panels it owns (one-panel look-ahead; csrc/sharded.cu).  The arithmetic per tile
 * is that of the single-GPU fit (same 512-column panels, same head kernel), so the factors agree bit for bit at any size.
 * When the call returns EVERY rank holds the complete factor and alpha, so predict / likelihood run locally (queries shard
 * trivially) and fgp_lml_gradient_sharded splits the O(n^3) inverse over the ranks.
 * fgp_comm_unique_id   rank 0: ncclGetUniqueId into a caller buffer of FGP_COMM_ID_BYTES bytes; the caller ships it to
 *                      the other processes (bench.py: torch.distributed broadcast — plumbing only).
 * fgp_comm_init_rank   every rank: ncclCommInitRank on the handle's GPU.  nranks == 1 needs no id exchange.
 * fgp_fit_sharded      fgp_fit, collectively. X / y_resid are read on rank 0 (others may pass NULL) and broadcast.
 * fgp_refit_sharded    fgp_refit, collectively (inputs already resident on every rank).
 * fgp_lml_gradient_sharded  fgp_lml_gradient, collectively (optimizer.rs:24-60, :159-203 on a sharded model): U = L^-T by
 *                      block rows r, r+P, .. (no communication), ncclAllGather of the rows, K^-1 = U U^T on the owned panels,
 *                      pair-tile reductions on the owned columns, one ncclAllReduce of the partial sums.  Same values on
 *                      every rank.
 * fgp_shard_plan       host-only: panel width, panel count, panels owned by `rank`, its share of the update flops.
 * fgp_comm_last_bytes  bytes this rank sent or received in panel broadcasts during the last sharded factorisation. */
#define FGP_COMM_ID_BYTES 128
int fgp_comm_unique_id(void* id_out, size_t bytes);
int fgp_comm_init_rank(fgp_model* m, const void* id, size_t bytes, int nranks, int rank);
int fgp_comm_destroy(fgp_model* m);
int fgp_fit_sharded(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
                    const fgp_kernel_desc* kernel, double noise, int has_eps, double eps);
int fgp_refit_sharded(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int has_eps, double eps);
int fgp_lml_gradient_sharded(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int scaled, double* scale_out,
                             double* grads);
int fgp_shard_plan(int64_t n, int nranks, int rank, int64_t* panel_cols, int64_t* n_panels, int64_t* n_owned,
                   double* flop_share);
double fgp_comm_last_bytes(const fgp_model* m);

/* measurement --------------------------------------------------------------------------------------------------
 * Device time (CUDA events on the model's stream) and number of kernel launches of the last entry-point call. */
double fgp_last_device_ms(const fgp_model* m);
int64_t fgp_last_launch_count(const fgp_model* m);
/* Per-kernel-class device timing of the last entry-point call (CUDA events around every launch on the model's stream).
 * Classes: 0 = gemm_nt (SYRK / GEMM / TRSM-as-GEMM, fp64 tensor pipe), 1 = potrf_diag / panel heads, 2 = pair tiles (Gram,
 * cross-covariance, gradient reductions), 3 = other, 4 = tcgen05 trailing updates (flops = f64-equivalent; the int8 tensor
 * work is 36 x that), 5 = digit slicing for them (flops slot = bytes moved).  ms / flops / count are arrays of 6. */
int fgp_set_profiling(fgp_model* m, int on);
/* Scheduling knobs (results are identical either way; used by bench.py / tests for A-B runs).
 * FGP_OPT_LOOKAHEAD (default 1): factor the next panel on a second, high-priority stream while the trailing update
 * of the current one runs. */
/* FGP_OPT_HEAD (default 1): factor each 512-column panel's diagonal block and its inverse in ONE multi-CTA launch
 * (csrc/potrf_head.cu) and solve everything below it as one K <= 512 GEMM; 0 = the per-block-column schedule
 * (diagonal tile / panel solve / rank-128 update launch triples). Results agree to rounding, not bit for bit. */
/* FGP_OPT_TCGEN05 (default 1): trailing updates behind a panel with at least 1024 rows left run on the 5th-generation
 * tensor cores (tcgen05.mma kind::i8 on exact base-128 digit slices of the panel, int32 accumulators in TMEM; csrc/ozaki.cuh);
 * 0 = the f64 DMMA kernel everywhere (A/B runs; both meet the 1e-10 factor tolerance). */
/* FGP_OPT_SHARD_PIPE (default -1 = automatic: 1 from 3 ranks up, else 0): 1 = in the multi-GPU fit every panel travels in row
 * pieces (solve, broadcast, digit slicing and the next owner's look-ahead overlap piece by piece; csrc/sharded.cu
 * factor_sharded_pipe); 0 = one piece per panel (factor_sharded_head; faster while the fit is work-bound, i.e. on 2 GPUs).
 * Same factor bit for bit.  The environment variable FGP_SHARD_PIPE=0 forces 0 for the whole process. */
enum fgp_option { FGP_OPT_LOOKAHEAD = 1, FGP_OPT_HEAD = 2, FGP_OPT_TCGEN05 = 3, FGP_OPT_SHARD_PIPE = 4 };
int fgp_set_option(fgp_model* m, int option, int64_t value);
int fgp_profile_summary(const fgp_model* m, double* ms, double* flops, int64_t* count);
/* LinearPrior::fit (src/parameters/prior.rs:139-159): least squares of the ORIGINAL outputs y (n values, host) on [1 | X] for the
 * resident training inputs. The normal equations are accumulated on the device over the centred inputs (the intercept
 * decouples), the d x d system is solved on the host; weights: d values. FGP_ERR_NOT_POSDEF when the inputs are rank deficient
 * (the reference's SVD solve would return a minimum-norm answer there); d <= 44. */
int fgp_linear_prior_fit(fgp_model* m, const double* y, double* weights, double* intercept);
/* Resident-input predict for kernel-only timing: stage queries once, then run the device part repeatedly. */
int fgp_stage_queries(fgp_model* m, const double* Xq, int64_t ldq, int64_t q);
int fgp_predict_staged(fgp_model* m, const fgp_kernel_desc* kernel, int want_mean, int want_var);
int fgp_fetch_predictions(fgp_model* m, double* mean_wo_prior, double* var);

/* page-locked host buffers for callers that want DMA-speed transfers (any host pointer is accepted everywhere) */
void* fgp_alloc_pinned(size_t bytes);
void fgp_free_pinned(void* p);

/* test hook: C = beta*C + alpha*A*B^T on device copies of host matrices, through the production GEMM kernel */
int fgp_dbg_gemm_nt(int device, double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                    int M, int N, int K, double alpha, int beta_one, int lower);

/* measurement hook: `reps` back-to-back launches of the production GEMM kernel on device-resident random matrices, SYRK-shaped
 * (C (M x N) -= A A[:N]^T, A is M x K; lower != 0: only tiles on/below the diagonal). *ms_out = CUDA-event time per launch,
 * *flops_out = algorithmic flops per launch. Used by tools/gemm_bench.py for the kernel's isolated roofline figure. */
int fgp_dbg_gemm_bench(int device, int M, int N, int K, int lower, int beta_one, int reps, double* ms_out, double* flops_out);

/* test hook: the panel head kernel (csrc/potrf_head.cu) alone: A is (128 nt)^2 column-major SPD, nt = 1..4 -> L in its lower
 * triangle; W (same shape, ld = 128 nt) <- L^-1; *info_out = 0, the 1-based failing column, or the kernel's time-out code */
int fgp_dbg_potrf_head(int device, double* A, int nt, double* W, int has_sub, double sub, int* info_out, int reps,
                       double* ms_out); /* reps > 0 and ms_out != NULL: also the kernel's CUDA-event time on fresh copies of A */

/* test hook, host only: the row pieces in which a panel with `below` rows under its diagonal block travels in the sharded fit
 * (csrc/sharded.cuh shard_pieces); returns their number (<= cap) or -1 */
int fgp_dbg_shard_pieces(int64_t below, int64_t pipe_rows, int64_t* first_row, int64_t* height, int cap);

/* test hook, host only: the branch-free exp(x), x <= 0, that the device kernels evaluate (csrc/kernel_eval.cuh exp_nonpos) */
double fgp_dbg_exp(double x);
/* test hook, host only: the table-assisted exp(x), x <= 0, of the Gram / cross-covariance interior tiles (exp_nonpos_tab:
 * 256-entry 2^(j/256) table + degree-4 polynomial; 0 for x <= -600) */
double fgp_dbg_exp_tab(double x);

/* test hook: resident CTAs per SM of the GEMM kernel on `device` (the design point is 2: one CTA's C read-modify-write
 * overlaps the other's DMMA main loop); -1 on error */
int fgp_dbg_gemm_occupancy(int device);
/* test hook, host only: rows per CTA (64 or 32) the GEMM launch of an M x N (lower != 0: triangular) problem uses on a GPU
 * with num_sms SMs — 32-row CTAs where they lower the heaviest SM's load (csrc/gemm_nt.cu gemm_nt_cta_rows) */
int fgp_dbg_gemm_cta_rows(int M, int N, int lower, int num_sms);
int fgp_dbg_gemm_occupancy32(int device); /* the 32-row shape used for sub-wave launches: 3 by design */

/* test hook, host only: the (tile row, tile column) each thread block of a lower-mode GEMM launch computes, for M x N
 * extents and tile-column groups of `grp` columns `stride` apart (the sharded trailing update); returns the tile count. */
int64_t fgp_dbg_lower_tiles(int M, int N, int grp, int stride, int* ti_out, int* tj_out, int64_t capacity);
/* the same with the first `row_skip` tile rows left out (the trailing update behind a panel whose successor's diagonal
 * block is updated by its own launch on the panel stream) */
int64_t fgp_dbg_lower_tiles_skip(int M, int N, int grp, int stride, int row_skip, int* ti_out, int* tj_out, int64_t capacity);

/* test hook: C (M x M, lower != 0: tiles on/below the diagonal, the first row_skip tile rows left out) -= A A^T, A is M x K
 * (K = 128..512), through the tcgen05 / TMEM exact-integer trailing update (csrc/ozaki.cu) on device copies of host
 * matrices; tiles_per_cta, lbo, sbo <= 0: production values */
int fgp_dbg_ozaki_syrk(int device, double* C, int64_t ldc, const double* A, int64_t lda, int M, int K, int lower, int row_skip,
                       int tiles_per_cta, int lbo, int sbo);
/* measurement hook: CUDA-event time per launch of the digit slicing kernel and of the tcgen05 update on random device data */
/* measurement switches of the update kernel: 1 = no epilogue arithmetic / stores, 2 = no operand copies, 4 = no MMAs; 0 = production */
void fgp_dbg_ozaki_experiment(int flags);
int fgp_dbg_ozaki_bench(int device, int M, int K, int reps, int tiles_per_cta, double* ms_update, double* ms_slice);

#ifdef __cplusplus
}
#endif
#endif
