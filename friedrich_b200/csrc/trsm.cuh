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

namespace fgp {

// returns the number of kernel launches
inline int64_t trsm_fwd_t(double* Xt, int64_t ldx, int64_t M, const double* L, int64_t ldl, const double* inv,
                          int64_t i_begin, int64_t i_end, double* trailing, const LaunchCtx& st, bool tri_rows = false) {
    int64_t launches = 0;
    for (int64_t i = i_begin; i < i_end; ++i) {
        const int64_t Mi = tri_rows ? std::min<int64_t>(M, (i + 1) * TILE) : M;
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

}  // namespace fgp
