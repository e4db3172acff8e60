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

constexpr int PAIR_TM = 128;  // tile rows
constexpr int PAIR_TN = 64;   // tile columns

// Epilogue concept:  __device__ void operator()(int64_t r, int64_t c, double dot, double d2);   (called for r>=c only
// when symmetric)    __device__ void finish();   (once per thread at the end; may use __syncthreads)
template <int MODE, class Epi>
__global__ void __launch_bounds__(256) pair_tile_kernel(PairArgs a, Epi epi) {
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

        // epilogue: fuse the norm broadcast and hand (dot, d2) to the functor
        double nrow[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) nrow[mi] = (MODE == PAIR_DOT) ? 0.0 : __ldg(a.na + wr0 + 8 * mi + g);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = wc0 + 8 * ni + 2 * t + e;
                const double ncol = (MODE == PAIR_DOT) ? 0.0 : __ldg(a.nb + c);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int64_t r = wr0 + 8 * mi + g;
                    if (a.symmetric && r < c) continue;
                    double dot = 0.0, d2 = 0.0;
                    if (MODE == PAIR_DOT) {
                        dot = acc[mi][ni][e];
                    } else {
                        if (MODE == PAIR_BOTH) dot = accr[mi][ni][e];
                        const double nsum = nrow[mi] + ncol;
                        d2 = fmax(nsum - 2.0 * acc[mi][ni][e], 0.0);
                        if (a.symmetric && r == c) {
                            d2 = 0.0;
                        } else if (d2 < nsum * 0x1p-12) {  // cancellation ate > 12 bits: direct differences
                            const double* xr = a.xa_c + r * dp;
                            const double* xc = a.xb_c + c * dp;
                            double s = 0.0;
                            for (int k = 0; k < dp; ++k) {
                                const double df = __ldg(xr + k) - __ldg(xc + k);
                                s = fma(df, df, s);
                            }
                            d2 = s;
                        }
                    }
                    epi(r, c, dot, d2);
                }
            }
        }
    }
    epi.finish();
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
template <int KIND>
struct CovWriteEpi {
    DevKernel k;
    double* out;
    int64_t ld;
    int64_t valid_rows, valid_cols;
    int symmetric;
    double noise2;
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, double dot, double d2) const {
        double v;
        if (r >= valid_rows || c >= valid_cols) {
            v = (symmetric && r == c) ? 1.0 : 0.0;
        } else {
            v = kernel_value<KIND>(k, dot, d2);
            if (symmetric && r == c) v += noise2;  // algebra/mod.rs:78
        }
        out[r + c * ld] = v;
    }
    __device__ __forceinline__ void finish() const {}
};

}  // namespace fgp
