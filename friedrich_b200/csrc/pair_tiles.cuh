// pair_tiles.cuh — tile engine for every "one kernel evaluation per (row point, column point)" matrix of the hot
// path: Gram lower triangle (algebra/mod.rs:70-79), cross-covariance (algebra/mod.rs:41-54), gradient Gram
// matrices consumed on the fly (algebra/mod.rs:129-155), mean pairwise distance (kernel.rs:94-113).
//
// The pair statistics come from the fp64 tensor pipe: dot(x_r, x_c) as a DMMA GEMM over the feature dimension,
// fused with the ||.||^2 broadcast (d2 = |x_r|^2 + |x_c|^2 - 2 x_r.x_c on CENTRED coordinates, so the cancellation
// error is relative to the data spread, not its offset) and an epilogue functor (kernel value / gradient / ...).
// Near-coincident pairs (d2 below 2^-12 of the norm sum, where the expansion has lost >12 bits) are recomputed by
// direct differences, and the diagonal of a symmetric matrix gets d2 = 0 exactly, as the reference's
// (x1 - x2).norm_squared() would (kernel.rs:558).
//
// Data layout: points are ROW-major [n_pad][dp] (dp = d rounded up to 4, zero filled) so a DMMA fragment load is a
// 32-byte contiguous read per point; the matrices written are column-major with leading dimension `ld`.
#pragma once

#include "common.cuh"
#include "kernel_eval.cuh"

namespace fgp {

enum PairMode { PAIR_D2 = 1, PAIR_DOT = 2, PAIR_BOTH = 3 };

struct PairArgs {
    const double* xa_c;  // row points, centred        [rows][dp]
    const double* xb_c;  // column points, centred     [cols][dp]
    const double* xa_r;  // row points, raw (PAIR_DOT / PAIR_BOTH)
    const double* xb_r;
    const double* na;    // |xa_c|^2 per row point
    const double* nb;    // |xb_c|^2 per column point
    int dp;
    int64_t rows, cols;  // padded extents (multiples of 128 / 64)
    int row_tile0;       // first 128-row tile visited (add_samples only rebuilds the last block rows)
    int col_tile0;       // first 64-column tile visited and how many (0 = all): the sharded fit assembles only the block
    int col_tiles;       // columns a rank owns
    int symmetric;       // rows and columns are the same point set: only r >= c is visited
};

// 2^(j/256) for exp_nonpos_tab (every translation unit that instantiates pair_tile_kernel keeps its own copy: no -rdc)
static __device__ const double g_exp2_tab[256] = {FGP_EXP2_TABLE_256};

constexpr int PAIR_TM = 128;  // tile rows
constexpr int PAIR_TN = 64;   // tile columns

// Epilogue concept:  __device__ void operator()(int64_t r, int64_t c, double dot, double d2);   (called for r>=c only
// when symmetric)    __device__ void skipped(int64_t r, int64_t c);   (the r<c elements of an active symmetric tile)
// __device__ void finish(int64_t row0, int64_t col0, bool active);   (once per thread at the end; may use __syncthreads)
// __device__ void bind(const Epi* param);   (the functor as passed to the kernel, in parameter space)
template <int MODE, class Epi>
__global__ void __launch_bounds__(256) pair_tile_kernel(PairArgs a, const __grid_constant__ Epi epi_in) {
    Epi epi = epi_in;   // accumulating epilogues keep per-thread state
    epi.bind(&epi_in);  // ... while a TMA descriptor must be addressed in the kernel-parameter space
    // exp table of the interior-tile fast path, scaled by the epilogue's amplitude (exp_nonpos_tab, kernel_eval.cuh)
    constexpr bool FAST = (MODE == PAIR_D2) && Epi::HAS_FAST;
    __shared__ double exp_tab[FAST ? 256 : 1];
    if (FAST) {
        exp_tab[threadIdx.x] = epi.fast_scale() * __ldg(&g_exp2_tab[threadIdx.x]);
        __syncthreads();
    }
    const int64_t row0 = (int64_t)(blockIdx.x + a.row_tile0) * PAIR_TM;
    const int64_t col0 = (int64_t)(blockIdx.y + a.col_tile0) * PAIR_TN;
    const bool active = !(a.symmetric && row0 + PAIR_TM - 1 < col0);
    if (active) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int g = lane >> 2, t = lane & 3;
        const int64_t wr0 = row0 + 32 * (warp & 3);
        const int64_t wc0 = col0 + 32 * (warp >> 2);
        const int dp = a.dp;

        double acc[4][4][2];
        double accr[(MODE == PAIR_BOTH) ? 4 : 1][(MODE == PAIR_BOTH) ? 4 : 1][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                if (MODE == PAIR_BOTH) accr[mi][ni][0] = accr[mi][ni][1] = 0.0;
            }

        const double* pa = ((MODE == PAIR_DOT) ? a.xa_r : a.xa_c) + (wr0 + g) * dp + t;
        const double* pb = ((MODE == PAIR_DOT) ? a.xb_r : a.xb_c) + (wc0 + g) * dp + t;
        for (int k0 = 0; k0 < dp; k0 += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) fa[mi] = __ldg(pa + (int64_t)(8 * mi) * dp + k0);
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) fb[ni] = __ldg(pb + (int64_t)(8 * ni) * dp + k0);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
        }
        if (MODE == PAIR_BOTH) {
            const double* qa = a.xa_r + (wr0 + g) * dp + t;
            const double* qb = a.xb_r + (wc0 + g) * dp + t;
            for (int k0 = 0; k0 < dp; k0 += 4) {
                double fa[4], fb[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) fa[mi] = __ldg(qa + (int64_t)(8 * mi) * dp + k0);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) fb[ni] = __ldg(qb + (int64_t)(8 * ni) * dp + k0);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(accr[mi][ni][0], accr[mi][ni][1], fa[mi], fb[ni]);
            }
        }

        // epilogue: fuse the norm broadcast and hand (dot, d2) to the functor.  Three passes so that the 32 kernel
        // evaluations of a thread form ONE branch-free block the scheduler can interleave: (1) d2 for every pair, (2) the
        // rare near-coincidence repair, (3) the functor.
        double nrow[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) nrow[mi] = (MODE == PAIR_DOT) ? 0.0 : __ldg(a.na + wr0 + 8 * mi + g);
        double d2v[4][4][2];
        unsigned fix = 0u;  // bit (ni*2+e)*4+mi: the expansion lost > 12 bits for this pair
        // INTERIOR tiles of a distance-kernel covariance matrix (no diagonal, no padding, nothing above the diagonal): the
        // per-pair guards (r == c, r < c, r / c beyond the valid extent) and the 64-bit index arithmetic of the general
        // path below are most of its instructions — here every pair is d2, the repair test, the kernel value and one store
        // at a 32-bit offset from a per-thread base.  (block-uniform branch)
        if (FAST && epi.fast_ok(row0, col0, a.symmetric)) {
            // the clamp at 0 and the repair test work on the high words (integer pipe; the fp64 pipe is this path's limit):
            // d2 < 0 <=> sign bit, d2 < nsum 2^-12 <=> exponent/mantissa-high comparison (both operands >= 0 there)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double ncol = __ldg(a.nb + wc0 + 8 * ni + 2 * t + e);
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const double nsum = nrow[mi] + ncol;
                        const double d2 = fma(-2.0, acc[mi][ni][e], nsum);
                        const int hd = __double2hiint(d2);
                        if (hd < __double2hiint(nsum) - (12 << 20)) fix |= 1u << ((ni * 2 + e) * 4 + mi);
                        d2v[mi][ni][e] = __hiloint2double(hd < 0 ? 0 : hd, hd < 0 ? 0 : __double2loint(d2));
                    }
                }
            if (fix) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                            if (fix & (1u << ((ni * 2 + e) * 4 + mi))) {
                                const double* xr = a.xa_c + (wr0 + 8 * mi + g) * dp;
                                const double* xc = a.xb_c + (wc0 + 8 * ni + 2 * t + e) * dp;
                                double sdiff = 0.0;
                                for (int k = 0; k < dp; ++k) {
                                    const double df = __ldg(xr + k) - __ldg(xc + k);
                                    sdiff = fma(df, df, sdiff);
                                }
                                d2v[mi][ni][e] = sdiff;
                            }
            }
            double* obase = epi.fast_base(wr0 + g, wc0 + 2 * t);
            const int ldi = epi.fast_ld();
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) obase[8 * mi + (8 * ni + e) * ldi] = epi.fast_value(d2v[mi][ni][e], exp_tab);
            epi.finish(row0, col0, active);
            return;
        }
        if (MODE != PAIR_DOT) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t c = wc0 + 8 * ni + 2 * t + e;
                    const double ncol = __ldg(a.nb + c);
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int64_t r = wr0 + 8 * mi + g;
                        const double nsum = nrow[mi] + ncol;
                        double d2 = fmax(nsum - 2.0 * acc[mi][ni][e], 0.0);
                        const bool diag = a.symmetric && r == c;
                        const bool skip = a.symmetric && r < c;
                        if (!diag && !skip && d2 < nsum * 0x1p-12) fix |= 1u << ((ni * 2 + e) * 4 + mi);
                        d2v[mi][ni][e] = diag ? 0.0 : d2;
                    }
                }
            if (fix) {  // cancellation ate > 12 bits: direct differences, as the reference's (x1 - x2).norm_squared()
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                            if (fix & (1u << ((ni * 2 + e) * 4 + mi))) {
                                const double* xr = a.xa_c + (wr0 + 8 * mi + g) * dp;
                                const double* xc = a.xb_c + (wc0 + 8 * ni + 2 * t + e) * dp;
                                double s = 0.0;
                                for (int k = 0; k < dp; ++k) {
                                    const double df = __ldg(xr + k) - __ldg(xc + k);
                                    s = fma(df, df, s);
                                }
                                d2v[mi][ni][e] = s;
                            }
            }
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = wc0 + 8 * ni + 2 * t + e;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int64_t r = wr0 + 8 * mi + g;
                    if (a.symmetric && r < c) {
                        epi.skipped(r, c);
                        continue;
                    }
                    const double dot = (MODE == PAIR_DOT) ? acc[mi][ni][e] : ((MODE == PAIR_BOTH) ? accr[mi][ni][e] : 0.0);
                    epi(r, c, dot, (MODE == PAIR_DOT) ? 0.0 : d2v[mi][ni][e]);
                }
            }
    }
    epi.finish(row0, col0, active);
}

inline dim3 pair_grid(const PairArgs& pa) {
    return dim3((unsigned)(pa.rows / PAIR_TM - pa.row_tile0),
                (unsigned)(pa.col_tiles > 0 ? pa.col_tiles : pa.cols / PAIR_TN - pa.col_tile0));
}

// Deterministic block reduction used by accumulating epilogues: sums NV per-thread values over the 256 threads of
// the CTA (fixed tree) and stores them at out[0..NV).
template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* out) {
    __shared__ double red[8][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue: write the covariance matrix.  symmetric => Gram lower triangle with noise^2 on the diagonal and an
// identity block on the padding (so the padded matrix stays positive definite and its factor is [[L,0],[0,I]]).
// Each lane group writes 64-byte column runs directly. (A shared-memory staged tile + one TMA tensor store measured 9 %
// slower: the kernel is bound by the ~100 instructions per pair — fp64 exp, the norm expansion and its guards — not by
// its stores; profiles/ncu_gram_*.jsonl.)
template <int KIND>
struct CovWriteEpi {
    DevKernel k;
    double* out;
    int64_t ld;
    int64_t valid_rows, valid_cols;
    int symmetric;
    double noise2;
    // host-precomputed constants of the two specialised kernels (one rounding away from the expressions as coded in
    // kernel.rs:556-560 / :873-878, which divide per pair): SquaredExp c0 = -1/(2 ls^2), c1 = |ampl|;
    // Matern2 c0 = sqrt(5)/|ls|, c1 = |ampl|, c2 = 5/(3 ls^2)
    double c0, c1, c2;
    __device__ __forceinline__ void bind(const CovWriteEpi*) {}
    // interior-tile fast path of pair_tile_kernel (the two specialised distance kernels only)
    static constexpr bool HAS_FAST = (KIND == KIND_SQEXP || KIND == KIND_MATERN2);
    __device__ __forceinline__ bool fast_ok(int64_t row0, int64_t col0, int sym) const {
        return ld < (int64_t)1 << 24 && row0 + PAIR_TM <= valid_rows && col0 + PAIR_TN <= valid_cols &&
               (!sym || row0 >= col0 + PAIR_TN) && c1 >= 0x1p-150 && c1 <= 0x1p150;  // amplitude range of exp_nonpos_tab
    }
    __device__ __forceinline__ double* fast_base(int64_t r, int64_t c) const { return out + r + c * ld; }
    __device__ __forceinline__ int fast_ld() const { return (int)ld; }
    __device__ __forceinline__ double fast_scale() const { return c1; }
    // tab = c1 * 2^(j/256): exp_nonpos_tab returns c1 * exp(.)
    __device__ __forceinline__ double fast_value(double d2, const double* tab) const {
        if (KIND == KIND_SQEXP) return exp_nonpos_tab(d2 * c0, tab);
        const double x = c0 * sqrt(d2);
        return (1.0 + x + c2 * d2) * exp_nonpos_tab(-x, tab);
    }
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, double dot, double d2) const {
        double v;
        if (KIND == KIND_SQEXP) v = c1 * exp_nonpos(d2 * c0);
        else if (KIND == KIND_MATERN2) {
            const double x = c0 * sqrt(d2);
            v = c1 * (1.0 + x + c2 * d2) * exp_nonpos(-x);
        } else v = kernel_value<KIND>(k, dot, d2);
        if (symmetric && r == c) v += noise2;  // algebra/mod.rs:78
        if (r >= valid_rows || c >= valid_cols) v = (symmetric && r == c) ? 1.0 : 0.0;
        out[r + c * ld] = v;
    }
    __device__ __forceinline__ void skipped(int64_t, int64_t) const {}
    __device__ __forceinline__ void finish(int64_t, int64_t, bool) const {}
};

}  // namespace fgp
