// potrf.cuh — blocked lower Cholesky on the padded device matrix (replaces nalgebra Cholesky::new /
// new_with_substitute called from make_cholesky_cov_matrix, src/algebra/mod.rs:81-91).
//
//   for each outer panel of panel_tiles(np) block columns:
//       for each 128-wide block column j of the panel:
//           L_jj = chol(A[j, j]),  inv_j = L_jj^-1                           potrf_diag_kernel (one CTA, see potrf.cu)
//           A[j+1:, j] = A[j+1:, j] * inv_j^T                                gemm_nt  (TRSM as a GEMM, in place)
//           A[c:, c] -= A[c:, j] * A[c, j]^T  for the panel's columns c > j   gemm_nt  (right-looking inside the panel, K = 128)
//       A[after:, after:] -= P * P^T,  P = A[after:, panel]                  gemm_nt  (SYRK, K = 128*panel_tiles)
//
// Schedule (potrf_lower): one-panel look-ahead — the next panel is updated and factored on a high-priority stream while the
// main stream applies the current panel to everything behind it; the look-ahead update itself is split over a third
// stream so that the next panel's first diagonal tile can start as soon as ITS block column is up to date.
//
// Failure semantics of the reference are kept: a pivot that is zero, negative or NaN is replaced by the
// substitute (`cholesky_epsilon`) when one is given and valid, otherwise the (1-based) failing column is
// recorded in `info` (first failure wins) and the host maps it to the reference's panic.
#pragma once

#include <functional>

#include "common.cuh"
#include "gemm_nt.cuh"

namespace fgp {

constexpr int DIAG_DS = 129;                                   // smem row stride of the 128x128 diagonal tile
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_SMEM_BYTES = (128 * DIAG_DS + 96 * 33 + 128) * 8;
// Outer panel width in 128-column tiles: 512 columns; 1024 from n = 24576 on when ONE GPU factors the matrix. A wider panel
// runs the trailing update at the GEMM kernel's better depth (K = 1024: 35.3 vs 34.4 TF/s) and halves its C traffic, but
// lengthens the serial panel chain, which only large single-GPU problems hide: measured n = 32768 360 -> 352 ms, n = 16384
// 50.3 -> 50.7 ms, and on 8 GPUs n = 32768 86 -> 102 ms (the chain is what bounds the sharded fit, and the block-cyclic
// balance coarsens), so the sharded schedule keeps 512. (A multi-GPU factor therefore equals the single-GPU one bit for bit
// below n = 24576 and to rounding, ~1e-14 normwise, above.)
inline int panel_tiles(int64_t np, int nranks = 1) { return (nranks == 1 && np >= 24576) ? 8 : 4; }

// ---- panel head (potrf_head.cu) ------------------------------------------------------------------------------------
constexpr int HEAD_PANEL = 512;   // columns of a panel = leading dimension of its inverse block W
constexpr int HEAD_SYNC_INTS = 32;
constexpr int HEAD_TIMEOUT = 0x7ffffff0;  // value of *info when a dependency wait inside the head kernel gave up  // counters one head launch uses (zeroed by the caller)
cudaError_t potrf_head_prepare();
int potrf_head_workers(int nt);     // worker CTAs of a head launch over nt block columns (grid = 1 + workers)
// Factor the (128 nt)^2 diagonal block at A (ld = lda) in place and form its inverse: inv[nt][128*128] (the diagonal tiles'
// inverses, zero upper triangle), W (512 x 512, ld 512: block (i, c), c <= i, of L11^-1), P = 512 x 512 scratch, sync =
// HEAD_SYNC_INTS zeroed ints.  col_base: global column of A(0,0) for the failure report.  ONE launch on c.st.
void launch_potrf_head(double* A, int64_t lda, int nt, double* inv, double* W, double* P, int* sync, int has_sub, double sub,
                       int* info, int col_base, const LaunchCtx& c);
// invT tiles = transposes of the nb inv tiles (the adjoint wavefront solve reads them); one launch
void launch_transpose_tiles(const double* inv, double* invT, int64_t nb, cudaStream_t st);

struct PotrfCounters {
    int64_t launches = 0;
};

// defined in potrf.cu
cudaError_t potrf_prepare();
// inv / invT of the nb diagonal tiles of an already factored matrix (restoring a serialised model, fgp_upload_state)
void launch_diag_inverse(const double* L, int64_t ld, int64_t nb, double* inv, double* invT, cudaStream_t st);
// Factor block columns [jb_begin, np/128) of the np x np matrix A (ld = lda) in place. Block columns before
// jb_begin must already hold final factor values in ALL rows (used by add_samples: the caller has applied them to
// the trailing block). invdiag / invdiagT: [np/128][128*128] (inverse blocks and their transposes). info: device int, 0 on entry.
// block columns [J, Jend) of one panel (left-looking inside the panel: update, diagonal tile, panel solve); launches on c.st
// Inside the panel the factorisation is right-looking in rank-128 steps: after block column j is final, the panel's remaining
// block columns get -= L[:, j] L[cols, j]^T. With a `side` stream only block column j+1 — the one the next diagonal tile waits
// for — is updated on c.st; block columns j+2.. are updated on the side stream, concurrently with that diagonal tile.
// `side->ev_join` must carry the side stream's last work on this panel's columns (the look-ahead update of block columns
// J+1.. by the previous panel) — c.st waits for it before it touches block column J+1. Every tile receives the same rank-128
// updates in the same order with or without a side stream, so the results are identical.
// `column_done` (optional): called on the host right after the launches that finalise block column j (diagonal tile + panel
// solve) have been enqueued on c.st — the sharded fit ships the column to the other GPUs from there.
struct PanelSide {
    cudaStream_t st;
    cudaEvent_t ev_fork, ev_join;
};
void factor_panel(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, double* invdiag, double* invdiagT,
                  int has_sub, double sub, int* info, const LaunchCtx& c, PotrfCounters* cnt, const PanelSide* side = nullptr,
                  const std::function<void(int64_t)>* column_done = nullptr);
// trailing block columns [c0, c1) (rows >= c0: the trapezoid on/below the diagonal) -= P P^T, P = block columns [J, Jend)
void trailing_update(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, int64_t c0, int64_t c1,
                     const LaunchCtx& c, PotrfCounters* cnt);

void trailing_update_cols(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, int64_t c0, int64_t c1,
                          const LaunchCtx& c, PotrfCounters* cnt);

// `la` (optional): a second, high-priority stream and two events for the one-panel look-ahead schedule; null => everything
// runs on st.st in program order.
struct PotrfLookahead {
    cudaStream_t panel;
    cudaEvent_t ev_panel, ev_trail;
    cudaStream_t side;   // optional: the look-ahead update of the next panel's block columns 1.. runs here, concurrently with
    cudaEvent_t ev_side; //           the factorisation of its block column 0 on `panel`; also the in-panel side updates
    cudaEvent_t ev_fork; //           (factor_panel's PanelSide = {side, ev_fork, ev_side})
};
// `after_first_panel_may_start` (optional, look-ahead schedule only): called on the host once the panel stream has been
// released to factor the first panel; whatever it enqueues on st.st (the Gram assembly of the columns BEHIND the first
// panel) runs concurrently with that factorisation and is waited for before anything else touches those columns.
void potrf_lower(double* A, int64_t lda, int64_t np, int64_t jb_begin, double* invdiag, double* invdiagT, int has_sub,
                 double sub, int* info, const LaunchCtx& st, const PotrfLookahead* la, PotrfCounters* cnt,
                 const std::function<void()>* after_first_panel_may_start = nullptr);

// ---- the head schedule (default) -------------------------------------------------------------------------------------
//   for each 512-column panel p = [J, Jend):
//       L11, W = L11^-1                 potrf_head_kernel on the diagonal block            (panel stream; ONE launch)
//       pbuf = A21 W^T                  gemm_nt, K <= 512, out of place (k_upto_col)        (top rows: panel stream; rest: main)
//       A(next diagonal block) -= ..    gemm_nt lower on the first nt2 tile rows            (panel stream) -> next head
//       A(everything else behind) -= pbuf pbuf^T     gemm_nt lower with row_skip           (main stream)
//       L21 <- pbuf                     2-D copy                                            (side stream)
// The chain between two heads is head -> solve of the next diagonal block's rows -> its update: three launches per 512
// columns instead of twelve, none of which needs a whole SM; everything full-height trails on the main stream.
struct PotrfWork {
    double* inv;      // [nb][128*128] inverse diagonal tiles
    double* invT;     // their transposes
    double* W;        // [panels][512*512] inverse of every panel's diagonal block (ld 512)
    double* P;        // 512*512 scratch of the head kernel
    int* sync;        // [panels][HEAD_SYNC_INTS]
    double* pbuf[2];  // >= (np - 128) * 512 doubles each: the solved panel below its diagonal block, ld = rows below
    cudaEvent_t ev_top, ev_rest, ev_copy[2];
    // tcgen05 trailing updates (csrc/ozaki.cuh): digit slices / row scales of the solved panel; null = f64 DMMA everywhere
    int8_t* oz_digits = nullptr;
    double* oz_scale = nullptr;
    // optional (full single-GPU fits): panel slot s keeps its digits at oz_digits + oz_off_bytes[s] / oz_scale + oz_off_rows[s]
    // (-1: the panel has fewer than OZ_MIN_ROWS rows below it and is never sliced); null: one scratch image, reused per panel
    const int64_t* oz_off_bytes = nullptr;
    const int64_t* oz_off_rows = nullptr;
};
// the trailing update behind a panel runs on tcgen05 when at least this many rows are left below the panel (a rule on the
// GLOBAL problem, so that the single-GPU and the sharded schedule treat every tile alike); below it the DMMA kernel's
// smaller work items win (measured: 2048 -> 1024 changes the n = 16384 fit within noise, n = 4096: 2.24 -> 2.20 ms, predict +17 %)
constexpr int64_t OZ_MIN_ROWS = 1024;
// `p0`: slot of the first panel's W / sync (panels are [jb_begin + 4 i, ..)); returns the number of panels factored.
int64_t potrf_lower_head(double* A, int64_t lda, int64_t np, int64_t jb_begin, const PotrfWork& w, int64_t p0, int has_sub,
                         double sub, int* info, const LaunchCtx& st, const PotrfLookahead* la, PotrfCounters* cnt,
                         const std::function<void()>* after_first_panel_may_start = nullptr);

}  // namespace fgp
