// gemm_nt.cuh — the fp64 tensor-pipe GEMM every O(n^3) step of the hot path is built from:
//
//        C(M x N) = beta * C + alpha * A(M x K) * B(N x K)^T        all column-major, "NT"
//
// Both operands have their non-contracted index contiguous in memory, which is how every product on the path is
// arranged (see DESIGN.md "everything is NT"): SYRK/GEMM trailing updates of the blocked Cholesky (replaces the
// axpy chain of nalgebra's Cholesky::new_internal, called at src/algebra/mod.rs:83,90), the panel solve
// L21 = A21 * inv(L11)^T, the multi-RHS forward solve on the transposed right-hand side, U = L^-T and K^-1 = U U^T.
//
// Execution model (one CTA per 128x128 output tile, 288 threads):
//   * warp 8  = producer: per k-chunk of 16 columns it arms an mbarrier and issues 32 TMA bulk copies (UBLKCP),
//     one 1 KiB column segment per lane, into a 4-stage shared-memory ring. Columns are laid out [k][132] doubles;
//     the 4-double pad makes the DMMA fragment reads (8 rows x 4 k per instruction) bank-conflict free.
//   * warps 0-7 = consumers: each owns a 32x64 sub-tile = 4x8 DMMA.8x8x4 accumulators (128 registers), waits on the
//     stage's "full" mbarrier, runs 4 k-steps x 32 DMMA, releases the stage on its "empty" mbarrier.
//   * epilogue: accumulators are combined with C in global memory (each lane group writes 64-byte column runs).
//
// All extents are multiples of the tile (matrices are padded), so there is no edge code.
#pragma once

#include "common.cuh"

namespace fgp {

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_KC = 16, GEMM_STAGES = 4;
constexpr int GEMM_LDS = GEMM_BM + 4;                                   // 132: smem column stride in doubles
constexpr int GEMM_STAGE_DOUBLES = 2 * GEMM_KC * GEMM_LDS;              // A + B tile of one stage
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_DOUBLES * 8 + 2 * GEMM_STAGES * 8;
constexpr int GEMM_THREADS = 288;

struct GemmArgs {
    double* C;
    int64_t ldc;
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    int M, N, K;     // multiples of 128, 128, 16
    double alpha;
    int beta_one;    // 1: C += alpha*A*B^T ; 0: C = alpha*A*B^T
    int lower;       // 1: only tiles on or below the diagonal are computed (N <= M: the first N/128 tile columns of the
                     //    triangle), diagonal tiles store r >= c only
    int grp, stride; // lower mode: the N/128 tile columns computed are groups of `grp` consecutive tile columns whose starts
                     //    are `stride` tile columns apart (0,0 = contiguous); rows/columns are relative to C, whose origin
                     //    is on the diagonal
    int k_from_tile; // 1: contraction starts at k = 128*max(tile_row, tile_col) (operands upper-triangular: U U^T)
};

// Lower mode: linear block index b -> tile (ti, tj), both relative to C (whose origin is on the diagonal).
// Tile columns come in groups of PT consecutive columns, group q starting at tile column q*S (S = PT: the plain triangle /
// trapezoid; S = P*PT: the panels one rank owns under the block-cyclic distribution).  Column tj holds the (tm - tj) tiles
// on or below the diagonal; blocks enumerate them column by column.
__host__ __device__ inline void lower_tile_decode(int tm, int PT, int S, int b, int& ti, int& tj) {
    const double a = (double)PT * tm - 0.5 * PT * (PT - 1), c2 = 0.5 * S * PT;
    const double disc = (a + c2) * (a + c2) - 4.0 * c2 * (double)b;
    int gi = (int)(((a + c2) - sqrt(disc > 0.0 ? disc : 0.0)) / (2.0 * c2));
    if (gi < 0) gi = 0;
    auto prefix = [&](int q) -> int64_t {  // tiles in groups 0 .. q-1
        return (int64_t)q * PT * tm - (int64_t)q * (PT * (PT - 1) / 2) - (int64_t)S * PT * ((int64_t)q * (q - 1) / 2);
    };
    while (gi > 0 && prefix(gi) > b) --gi;
    while (tm - (int64_t)(gi + 1) * S > 0 && prefix(gi + 1) <= b) ++gi;
    int rem = b - (int)prefix(gi);
    int w = 0;
    for (; w < PT - 1; ++w) {
        const int cnt = tm - gi * S - w;
        if (rem < cnt) break;
        rem -= cnt;
    }
    tj = gi * S + w;
    ti = tj + rem;
}

// defined in gemm_nt.cu
int64_t gemm_nt_tiles(const GemmArgs& g);  // number of 128x128 tiles one launch computes
cudaError_t gemm_nt_prepare();
// algorithmic flops of one launch (what the roofline figure in bench.py is computed from)
double gemm_nt_flops(const GemmArgs& g);
// launches nothing when the problem is empty; returns the number of tiles launched
int64_t gemm_nt_launch(const GemmArgs& g, const LaunchCtx& ctx);

}  // namespace fgp
