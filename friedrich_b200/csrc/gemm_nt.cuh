// gemm_nt.cuh — the fp64 tensor-pipe GEMM every O(n^3) step of the hot path is built from:
//
//        C(M x N) = beta * C + alpha * A(M x K) * B(N x K)^T        all column-major, "NT"
//
// Both operands have their non-contracted index contiguous in memory, which is how every product on the path is
// arranged (see DESIGN.md "everything is NT"): SYRK/GEMM trailing updates of the blocked Cholesky (replaces the
// axpy chain of nalgebra's Cholesky::new_internal, called at src/algebra/mod.rs:83,90), the panel solve
// L21 = A21 * inv(L11)^T, the multi-RHS forward solve on the transposed right-hand side, U = L^-T and K^-1 = U U^T.
//
// Execution model (CTAs of 128 threads, two resident per SM; a CTA computes one 64 x 128 half tile — the two row halves of a
// 128x128 tile are consecutive blocks; 8 warps per SM is the most the register file allows at 128 accumulator registers
// per thread: 16 K registers per SM sub-partition / 2 warps):
//   * every warp owns a 32x64 sub-tile = 4x8 DMMA.8x8x4 accumulators (128 registers); per k-chunk of 16 columns it waits on
//     the stage's "full" mbarrier, runs 4 k-steps x 32 DMMA, and releases the stage on its "empty" mbarrier;
//   * there is no producer warp: the warps take turns (chunk index mod 4) refilling the 4-stage shared-memory ring two
//     chunks ahead — wait for the stage's "empty" barrier, arm "full", and one lane issues two TMA tensor copies (UTMALDG):
//     a (64+4) x 16 box of A and a (128+4) x 16 box of B. The 4 extra rows per column are never read: they ARE the padding
//     that makes the shared-memory column stride 68 / 132 doubles, which keeps the DMMA fragment reads (8 rows x 4 k per
//     instruction) bank-conflict free. (Per-lane 1-D bulk copies, one column each, cost 7 % of the kernel: the 32 UBLKCP
//     of a chunk are serialised through uniform registers.)
//   * epilogue: alpha * acc is staged in the idle ring as a dense 64 x 128 image and leaves in ONE TMA tensor operation —
//     a reduce-add into C (UTMAREDG; the f64 add is performed at the L2, the SM never reads C) or a plain store. While
//     one CTA of the SM is in its prologue / epilogue the other keeps the DMMA pipe busy;
//   * lower-mode launches walk the triangle in bands of GEMM band rows (tile rows), column by column inside a band, so
//     that the A blocks of a band stay L2-resident while the B blocks stream (gemm_nt_plan / gemm_tile_decode).
//   * sub-wave launches whose heaviest SM gets lighter that way use a second instantiation with 32-row CTAs (four per tile, 1 x 4
//     warps of 32 x 32, 3-stage ring, three resident per SM; `GemmShape` in gemm_nt.cu): twice as many, half as long CTAs
//     spread the latency-bound panel products evenly over the SMs. Both shapes sum every element in the same order.
//   Splitting the tile by ROWS keeps the in-place products (C aliases A: panel solve, multi-RHS solve) race free: a CTA
//   only ever reads the rows of A it later overwrites. In lower mode the strict upper triangle of a diagonal tile is left
//   unchanged by beta = 1 launches and zeroed by beta = 0 launches.
//
// All extents are multiples of the tile (matrices are padded), so there is no edge code.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace fgp {

// tools/microbench/gemm_variants.cu sweeps the ring geometry through these macros; production uses the defaults
#ifndef FGP_GEMM_KC
#define FGP_GEMM_KC 16
#define FGP_GEMM_STAGES 4
#define FGP_GEMM_AHEAD 2
#endif
constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_KC = FGP_GEMM_KC, GEMM_STAGES = FGP_GEMM_STAGES;  // BM x BN: the tile the grid counts
constexpr int GEMM_CTA_M = 64;                                          // rows of the tile one CTA computes
constexpr int GEMM_LDA = GEMM_CTA_M + 4;                                // 68: smem column stride of the A stage (doubles)
constexpr int GEMM_LDB = GEMM_BN + 4;                                   // 132: smem column stride of the B stage
constexpr int GEMM_STAGE_DOUBLES = GEMM_KC * (GEMM_LDA + GEMM_LDB);     // A + B tile of one stage
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_DOUBLES * 8 + 2 * GEMM_STAGES * 8;
constexpr int GEMM_WARPS = 4;
constexpr int GEMM_AHEAD = FGP_GEMM_AHEAD;                                           // k-chunks in flight ahead of the one consumed
constexpr int GEMM_THREADS = 32 * GEMM_WARPS;
constexpr int GEMM_MAX_BANDS = 64, GEMM_BAND_ROWS = 16;

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
    int k_tile0;     // k_from_tile: tile index of C's origin on the diagonal of the full matrix (the contraction of tile (ti, tj)
                     //    starts at k = 128*(k_tile0 + max(ti, tj)): a rank's share of K^-1 = U U^T starts at its first owned panel)
    int k_upto_col;  // 1: contraction ends at k = 128*(tile_col + 1) (B lower block-triangular: the panel solve A W^T, W = L11^-1)
    int row_skip;    // lower mode: the first row_skip tile rows are not computed (the diagonal block of the next panel is
                     //    updated by a separate, earlier launch on the panel stream)
    // filled by gemm_nt_plan (called by gemm_nt_launch): the band rasterisation of lower-mode launches
    int band_rows;                          // tile rows per band
    int n_bands;
    int band_prefix[GEMM_MAX_BANDS + 1];    // tiles in bands 0 .. r-1
    // filled by gemm_nt_launch: blocks [stagger_lo, stagger_hi) — the second CTA each SM receives in the first wave — start
    // stagger_ns late, so that the two CTAs of an SM stay half a tile out of phase for the rest of the launch and one's
    // prologue / epilogue falls into the other's main loop (0 = off: launches of only a few waves)
    int stagger_lo, stagger_hi, stagger_ns;
};

// Lower mode: the tiles of the launch, relative to C (whose origin is on the diagonal): local tile column jl sits at tile
// column tj = (jl / PT) * S + jl % PT (groups of PT consecutive columns, group q starting at tile column q*S; S = PT: the
// plain triangle / trapezoid; S = P*PT: the panels one rank owns under the block-cyclic distribution) and holds the tiles
// ti = tj .. tm-1.  They are enumerated band by band (band r = tile rows [r*R, (r+1)*R)), inside a band column by column,
// inside a column top to bottom.  gemm_tile_decode maps the linear tile index to (ti, tj) using the per-band prefix counts.
__host__ __device__ inline void gemm_tile_decode(const GemmArgs& g, int b, int& ti, int& tj) {
    const int tm = g.M / GEMM_BM;
    if (!g.lower) {
        ti = b % tm;
        tj = b / tm;
        // A W^T with block-triangular W: tile column j contracts over 128 (j + 1) columns — heaviest tile columns first, so that
        // the last wave of the launch is made of the light ones
        if (g.k_upto_col) tj = g.N / GEMM_BN - 1 - tj;
        return;
    }
    const int PT = g.grp > 0 ? g.grp : 1, S = g.stride > 0 ? g.stride : 1, tn = g.N / GEMM_BN, R = g.band_rows;
    int r = 0;
    while (r + 1 < g.n_bands && g.band_prefix[r + 1] <= b) ++r;
    int o = b - g.band_prefix[r];
    const int lo = g.row_skip + r * R, hi = (lo + R < tm) ? lo + R : tm, h = hi - lo;
    // local columns entirely above the band (tj < lo) are full-height
    int nf = (lo / S) * PT + ((lo % S) < PT ? (lo % S) : PT);
    if (nf > tn) nf = tn;
    if (o < nf * h) {
        const int jl = o / h;
        tj = (jl / PT) * S + jl % PT;
        ti = lo + o % h;
        return;
    }
    o -= nf * h;
    int jl = nf;
    for (;; ++jl) {
        tj = (jl / PT) * S + jl % PT;
        const int cnt = hi - tj;
        if (o < cnt) break;
        o -= cnt;
    }
    ti = tj + o;
}

// defined in gemm_nt.cu
int64_t gemm_nt_tiles(const GemmArgs& g);  // number of 128x128 tiles one launch computes
int gemm_nt_cta_rows(int64_t tiles, int num_sms);  // 64 or 32: the CTA shape a launch of `tiles` tiles uses (host rule)
void gemm_nt_plan(GemmArgs& g);             // fills band_rows / n_bands / band_prefix (host)
cudaError_t gemm_nt_prepare();
// TMA descriptor of a rows x cols column-major f64 matrix (leading dimension ld) moved in box_rows x box_cols tiles
// (cuTensorMapEncodeTiled through the runtime's driver entry point; false on failure)
bool make_tile_map(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
void gemm_nt_flag_error();   // a kernel launch could not be set up (tensor-map encode failure): reported by end_timed
bool gemm_nt_take_error();  // true once after a launch on the CURRENT device could not be set up (tensor-map encode failure)
int gemm_nt_num_sms();      // SM count of the current device (cached per device)
int gemm_nt_occupancy(int cta_rows = 64);  // resident CTAs per SM of the 64-row (2 by design) / 32-row (3) kernel, -1 on error
// algorithmic flops of one launch (what the roofline figure in bench.py is computed from)
double gemm_nt_flops(const GemmArgs& g);
// launches nothing when the problem is empty; returns the number of tiles launched
int64_t gemm_nt_launch(const GemmArgs& g, const LaunchCtx& ctx);

}  // namespace fgp
