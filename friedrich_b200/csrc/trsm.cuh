// trsm.cuh — multi-right-hand-side forward substitution with the Cholesky factor, in the TRANSPOSED layout the
// tensor-pipe GEMM wants:  Xt (M x n, column-major)  <-  Xt * L^-T      (i.e. X = L^-1 B with Xt = X^T, B = Xt^T).
//
// Replaces nalgebra's column-oriented `solve_lower_triangular` on an n x q right-hand side
// (src/gaussian_process/mod.rs:203, :260-263, :342-345 and the first half of Cholesky::solve, :235, :298, :379) and the
// `L11 r = c` solve inside `insert_column` (src/algebra/mod.rs:124).
//
// Right-looking over 128-wide block columns, with the diagonal blocks already inverted by potrf_diag_kernel, so every
// step is an NT GEMM on the fp64 tensor pipe:
//       Xt[:, i]      = Xt[:, i] * inv_i^T                      (M x 128 x 128, in place)
//       Xt[:, i+1:]  -= Xt[:, i] * L[i+1:, i]^T                 (M x (n - 128(i+1)) x 128)
// `tri_rows`: the right-hand side is the identity (U = L^-T for the explicit inverse, optimizer.rs:32,169): block
// column i of Xt is non-zero only in its first 128(i+1) rows, so M shrinks accordingly (n^3/3 instead of n^3 flops).
// `trailing`: add_samples — after the solve, trailing (M x M, lower) -= Xt * Xt^T in one SYRK over all block columns.
#pragma once

#include "gemm_nt.cuh"
#include "ozaki.cuh"

namespace fgp {

// returns the number of kernel launches
// `tri_first` / `tri_stride` (with tri_rows): the rows of Xt are the identity rows of the block rows tri_first, tri_first +
// tri_stride, ... only (a rank's share of U = L^-T in the sharded LML gradient): block column i is then non-zero in the first
// 128 * #{owned block rows <= i} rows.
inline int64_t trsm_fwd_t(double* Xt, int64_t ldx, int64_t M, const double* L, int64_t ldl, const double* inv,
                          int64_t i_begin, int64_t i_end, double* trailing, const LaunchCtx& st, bool tri_rows = false,
                          int64_t tri_first = 0, int64_t tri_stride = 1) {
    int64_t launches = 0;
    for (int64_t i = i_begin; i < i_end; ++i) {
        const int64_t Mi = tri_rows ? std::min<int64_t>(M, (i >= tri_first ? (i - tri_first) / tri_stride + 1 : 0) * TILE) : M;
        if (Mi <= 0) continue;
        {
            GemmArgs g{};
            g.C = Xt + i * TILE * ldx; g.ldc = ldx;
            g.A = g.C; g.lda = ldx;
            g.B = inv + i * TILE * TILE; g.ldb = TILE;
            g.M = (int)Mi; g.N = TILE; g.K = TILE;
            g.alpha = 1.0; g.beta_one = 0; g.lower = 0; g.k_from_tile = 0;
            launches += gemm_nt_launch(g, st) > 0;
        }
        if (i + 1 < i_end) {
            GemmArgs g{};
            g.C = Xt + (i + 1) * TILE * ldx; g.ldc = ldx;
            g.A = Xt + i * TILE * ldx; g.lda = ldx;
            g.B = L + (i + 1) * TILE + i * TILE * ldl; g.ldb = ldl;
            g.M = (int)Mi; g.N = (int)((i_end - i - 1) * TILE); g.K = TILE;
            g.alpha = -1.0; g.beta_one = 1; g.lower = 0; g.k_from_tile = 0;
            launches += gemm_nt_launch(g, st) > 0;
        }
    }
    if (trailing && i_end > i_begin) {
        GemmArgs g{};
        g.C = trailing; g.ldc = ldx;
        g.A = Xt + i_begin * TILE * ldx; g.lda = ldx;
        g.B = g.A; g.ldb = ldx;
        g.M = g.N = (int)M; g.K = (int)((i_end - i_begin) * TILE);
        g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.k_from_tile = 0;
        launches += gemm_nt_launch(g, st) > 0;
    }
    return launches;
}

// The same solve in 512-column panels with a one-panel look-ahead on a second stream (the schedule of potrf_lower):
//   panel stream   inside the panel: Xt[:, i] *= inv_i^T, Xt[:, i+1 .. panel end) -= Xt[:, i] L[.., i]^T   (K = 128, short)
//                  then the NEXT panel's columns -= Xt[:, panel] L[next, panel]^T                           (K = 512)
//   main stream    all later columns            -= Xt[:, panel] L[later, panel]^T                           (K = 512)
// so the 128-long chain of small in-panel products hides behind the large K = 512 updates, which also run the GEMM kernel at
// its efficient depth.  Every product sees the operands of the sequential algorithm; only the grouping of the k-sums
// differs (4 blocks of 128 accumulated in one launch).  Returns the number of launches; ends joined on st.st.
inline int64_t trsm_fwd_t_lookahead(double* Xt, int64_t ldx, int64_t M, const double* L, int64_t ldl, const double* inv,
                                    int64_t nb, const LaunchCtx& st, cudaStream_t panel_stream, cudaEvent_t ev_panel,
                                    cudaEvent_t ev_trail) {
    constexpr int64_t PT = 4;
    int64_t launches = 0;
    LaunchCtx pc = st;
    pc.st = panel_stream;
    auto gemm = [&](double* C, const double* A, const double* B, int64_t ldb, int64_t N, int64_t K, bool solve,
                    const LaunchCtx& c) {
        GemmArgs g{};
        g.C = C; g.ldc = ldx;
        g.A = A; g.lda = ldx;
        g.B = B; g.ldb = ldb;
        g.M = (int)M; g.N = (int)N; g.K = (int)K;
        g.alpha = solve ? 1.0 : -1.0; g.beta_one = solve ? 0 : 1; g.lower = 0; g.k_from_tile = 0;
        launches += gemm_nt_launch(g, c) > 0;
    };
    auto inner = [&](int64_t P, int64_t Pend) {
        for (int64_t i = P; i < Pend; ++i) {
            double* Xi = Xt + i * TILE * ldx;
            gemm(Xi, Xi, inv + i * TILE * TILE, TILE, TILE, TILE, true, pc);
            if (i + 1 < Pend) gemm(Xi + TILE * ldx, Xi, L + (i + 1) * TILE + i * TILE * ldl, ldl, (Pend - i - 1) * TILE, TILE, false, pc);
        }
    };
    cudaEventRecord(ev_trail, st.st);  // the panel stream starts after whatever filled Xt on the main stream
    cudaStreamWaitEvent(panel_stream, ev_trail, 0);
    bool first = true;
    for (int64_t P = 0; P < nb; P += PT) {
        const int64_t Pend = std::min(P + PT, nb);
        inner(P, Pend);
        cudaEventRecord(ev_panel, panel_stream);
        if (Pend < nb) {
            const int64_t Pend2 = std::min(Pend + PT, nb);
            const double* Xp = Xt + P * TILE * ldx;
            cudaStreamWaitEvent(st.st, ev_panel, 0);                          // main: panel [P, Pend) is solved
            if (!first) cudaStreamWaitEvent(panel_stream, ev_trail, 0);       // panel: earlier updates reached columns >= Pend
            gemm(Xt + Pend * TILE * ldx, Xp, L + Pend * TILE + P * TILE * ldl, ldl, (Pend2 - Pend) * TILE, (Pend - P) * TILE,
                 false, pc);
            if (Pend2 < nb)
                gemm(Xt + Pend2 * TILE * ldx, Xp, L + Pend2 * TILE + P * TILE * ldl, ldl, (nb - Pend2) * TILE,
                     (Pend - P) * TILE, false, st);
            cudaEventRecord(ev_trail, st.st);
            first = false;
        }
    }
    cudaStreamWaitEvent(st.st, ev_panel, 0);  // join
    return launches;
}

// The same solve over the PANELS of the head schedule, with the inverse W_p = L11^-1 of every panel's diagonal block
// (potrf.cuh PotrfWork::W): per panel one out-of-place product  T = Xt[:, panel] W_p^T  (K <= 512, block-triangular W) and
// one update  Xt[:, later] -= T L[later, panel]^T  (K = panel width): 2 launches per 512 columns instead of 8, all of them at
// the GEMM kernel's efficient depth.  `tmp0/1` hold M x 512 doubles (ld = M); panel p covers block columns
// [pstart[p], pstart[p+1]) (pstart[npanels] = i_end).  With a panel stream the NEXT panel's columns are updated first on it and
// its T product follows at once, while the main stream updates everything behind (one-panel look-ahead as in potrf_lower_head).
// `upd_end` >= i_end: block columns [i_end, upd_end) are updated like the others but never solved — add_samples: the rows of
// Xt ARE the new block rows of L, so L[later, panel] for those columns is the panel just solved (copied back before the
// update reads it) and the new diagonal block receives  -= T T^T  panel by panel, in order (no K = n SYRK at the end).
// Returns the number of launches; ends joined on st.st.
// `ozs` (optional): digit slices of L[below, panel] kept by the fit (model.cuh ozL), covering block columns < i_end.  The main
// update of a panel that has them runs on tcgen05 (csrc/ozaki.cuh): T is sliced into the scratch image, then
// Xt[:, later] -= T L[later, panel]^T as exact int8 products; panels without digits (fewer than OZ_MIN_ROWS rows below) and
// the next panel's columns on the panel stream keep the f64 DMMA kernel.
struct OzPanelStore {
    const int8_t* digits;
    const double* scale;
    const int64_t* off_bytes;   // per panel, -1 = none
    const int64_t* off_rows;
    int8_t* scratch;            // >= M x 512 x 8 bytes
    double* scratch_scale;      // >= M doubles
};
inline int64_t trsm_fwd_t_panels(double* Xt, int64_t ldx, int64_t M, const double* L, int64_t ldl, const double* W,
                                 const int64_t* pstart, int64_t npanels, int64_t i_end, int64_t upd_end, double* tmp0,
                                 double* tmp1, const LaunchCtx& st, cudaStream_t panel_stream, cudaEvent_t ev_panel,
                                 cudaEvent_t ev_trail0, cudaEvent_t ev_trail1, const OzPanelStore* ozs = nullptr,
                                 bool tri_rows = false) {
    // `tri_rows`: the right-hand side is the identity (U = L^-T): the columns of panel [J, Jend) are non-zero in the first
    // 128 Jend rows only, so every product of the panel works on Mi = min(M, 128 Jend) rows (n^3/3 instead of n^3 flops)
    constexpr int64_t WP = 512;
    int64_t launches = 0;
    LaunchCtx pc = st;
    const bool two = panel_stream != nullptr;
    if (two) pc.st = panel_stream;
    int64_t Mi = M;   // rows the current panel's products work on
    auto gemm = [&](double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t N, int64_t K,
                    double alpha, int beta_one, int k_upto, const LaunchCtx& c) {
        if (N <= 0) return;
        GemmArgs g{};
        g.C = C; g.ldc = ldc;
        g.A = A; g.lda = lda;
        g.B = B; g.ldb = ldb;
        g.M = (int)Mi; g.N = (int)N; g.K = (int)K;
        g.alpha = alpha; g.beta_one = beta_one; g.lower = 0; g.k_upto_col = k_upto;
        launches += gemm_nt_launch(g, c) > 0;
    };
    auto pend = [&](int64_t p) { return p + 1 < npanels ? pstart[p + 1] : i_end; };
    if (upd_end < i_end) upd_end = i_end;
    // Xt[:, block columns [c0, upd_end)] -= T L[those rows, panel p]^T on the main stream: tcgen05 when the panel's digits exist
    auto update_main = [&](int64_t p, int64_t J, int64_t Jend, int64_t c0, int64_t w, const double* tmp) {
        if (c0 >= upd_end) return;
        const int64_t c_oz = std::min(i_end, upd_end);   // the kept digits cover the columns [.., i_end) (rows of L that existed at the fit)
        if (ozs && ozs->off_bytes[p] >= 0 && c0 < c_oz) {
            if (upd_end > c_oz)   // add_samples: the new block's own columns read the panel just solved (copied back): f64 DMMA
                gemm(Xt + c_oz * TILE * ldx, ldx, tmp, M, L + c_oz * TILE + J * TILE * ldl, ldl, (upd_end - c_oz) * TILE, w, -1.0, 1, 0, st);
            ozaki_slice_launch(tmp, M, Mi, (int)w, ozs->scratch, ozs->scratch_scale, st);
            GemmArgs g{};
            g.C = Xt + c0 * TILE * ldx; g.ldc = ldx;
            g.M = (int)Mi; g.N = (int)((c_oz - c0) * TILE); g.K = (int)w;
            g.alpha = -1.0; g.beta_one = 1; g.lower = 0;
            const int64_t toff = c0 - Jend;   // the panel's digit image starts at the first row below its diagonal block
            launches += 2 + (ozaki_update_launch(g, ozs->scratch, ozs->scratch_scale,
                                                 ozs->digits + ozs->off_bytes[p] + toff * (w / OZ_KSTEP) * (int64_t)OZ_PART_BYTES,
                                                 ozs->scale + ozs->off_rows[p] + toff * TILE, 0, st) > 0);
        } else {
            gemm(Xt + c0 * TILE * ldx, ldx, tmp, M, L + c0 * TILE + J * TILE * ldl, ldl, (upd_end - c0) * TILE, w, -1.0, 1, 0, st);
        }
    };
    if (two) {
        cudaEventRecord(ev_trail0, st.st);  // the panel stream starts after whatever filled Xt on the main stream
        cudaStreamWaitEvent(panel_stream, ev_trail0, 0);
    }
    for (int64_t p = 0; p < npanels; ++p) {
        const int64_t J = pstart[p], Jend = pend(p), w = (Jend - J) * TILE;
        double* tmp = (p & 1) ? tmp1 : tmp0;
        cudaEvent_t ev_trail = (p & 1) ? ev_trail1 : ev_trail0;
        double* Xp = Xt + J * TILE * ldx;
        Mi = tri_rows ? std::min<int64_t>(M, Jend * TILE) : M;
        // main-stream update p-2 has read tmp of this parity and brought this panel's columns up to date
        if (two && p >= 2) cudaStreamWaitEvent(panel_stream, ev_trail, 0);
        // T = Xt[:, panel] W_p^T ; the solved panel goes back into Xt (a small 2-D copy)
        gemm(tmp, M, Xp, ldx, W + p * WP * WP, WP, w, w, 1.0, 0, 1, pc);
        cudaMemcpy2DAsync(Xp, ldx * sizeof(double), tmp, M * sizeof(double), Mi * sizeof(double), (size_t)w, cudaMemcpyDeviceToDevice,
                          pc.st);
        if (Jend >= upd_end) break;
        // the next panel's columns first (panel stream: its T product follows at once), everything behind them on the main stream
        const int64_t Jnext = (Jend < i_end) ? pend(p + 1) : Jend;
        if (two) {
            cudaEventRecord(ev_panel, panel_stream);
            // the next panel's columns also receive the main-stream update of panel p-1 (they lie behind ITS next panel): wait for
            // it, so that every element gets its additions in panel order whatever the timing (results reproducible bit for bit)
            if (p >= 1) cudaStreamWaitEvent(panel_stream, (p & 1) ? ev_trail0 : ev_trail1, 0);
            gemm(Xt + Jend * TILE * ldx, ldx, tmp, M, L + Jend * TILE + J * TILE * ldl, ldl, (Jnext - Jend) * TILE, w, -1.0, 1, 0, pc);
            cudaStreamWaitEvent(st.st, ev_panel, 0);
            update_main(p, J, Jend, Jnext, w, tmp);
            cudaEventRecord(ev_trail, st.st);
        } else {
            update_main(p, J, Jend, Jend, w, tmp);
        }
    }
    if (two) {
        cudaEventRecord(ev_panel, panel_stream);
        cudaStreamWaitEvent(st.st, ev_panel, 0);  // join
    }
    return launches;
}

}  // namespace fgp
