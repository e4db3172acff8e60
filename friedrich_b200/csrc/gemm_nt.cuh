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
    int lower;       // 1: only tiles on or below the diagonal are computed, diagonal tiles store r >= c only
    int k_from_tile; // 1: contraction starts at k = 128*max(tile_row, tile_col) (operands upper-triangular: U U^T)
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nt_kernel(GemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GEMM_STAGES * GEMM_STAGE_DOUBLES * 8);
    uint64_t* empty = full + GEMM_STAGES;

    // tile coordinates: linear block index -> (ti, tj); lower mode enumerates the triangle column by column
    int ti, tj;
    {
        const int tm = g.M / GEMM_BM;
        int b = blockIdx.x;
        if (g.lower) {
            // column tj holds (tm - tj) tiles; find tj with prefix(tj) <= b < prefix(tj+1), prefix(j) = j*tm - j(j-1)/2
            double tmf = (double)tm + 0.5;
            int j = (int)(tmf - sqrt(tmf * tmf - 2.0 * (double)b));
            if (j < 0) j = 0;
            while (j > 0 && (int64_t)j * tm - (int64_t)j * (j - 1) / 2 > b) --j;
            while ((int64_t)(j + 1) * tm - (int64_t)(j + 1) * j / 2 <= b) ++j;
            tj = j;
            ti = j + (b - (int)((int64_t)j * tm - (int64_t)j * (j - 1) / 2));
        } else {
            ti = b % tm;
            tj = b / tm;
        }
    }
    const int m0 = ti * GEMM_BM, n0 = tj * GEMM_BN;
    const int kbeg = g.k_from_tile ? GEMM_BM * max(ti, tj) : 0;
    const int nk = (g.K - kbeg) / GEMM_KC;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 8) {
        // ---------------- producer ----------------
        const double* src = (lane < 16) ? g.A + m0 + (int64_t)(kbeg + lane) * g.lda
                                        : g.B + n0 + (int64_t)(kbeg + lane - 16) * g.ldb;
        const int64_t step = (int64_t)GEMM_KC * ((lane < 16) ? g.lda : g.ldb);
        const int dst_off = ((lane < 16) ? lane : (GEMM_KC + lane - 16)) * GEMM_LDS;
        for (int it = 0; it < nk; ++it) {
            const int s = it % GEMM_STAGES;
            if (it >= GEMM_STAGES) mbar_wait(&empty[s], ((it / GEMM_STAGES) + 1) & 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], 2 * GEMM_KC * GEMM_BM * 8);
            __syncwarp();
            tma_load_1d(tiles + (size_t)s * GEMM_STAGE_DOUBLES + dst_off, src, GEMM_BM * 8, &full[s]);
            src += step;
        }
    } else {
        // ---------------- consumers ----------------
        const int gq = lane >> 2, t = lane & 3;
        const int wr = warp & 3, wc = warp >> 2;
        double acc[4][8][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int it = 0; it < nk; ++it) {
            const int s = it % GEMM_STAGES;
            mbar_wait(&full[s], (it / GEMM_STAGES) & 1);
            const double* As = tiles + (size_t)s * GEMM_STAGE_DOUBLES + t * GEMM_LDS + 32 * wr + gq;
            const double* Bs = As - 32 * wr + GEMM_KC * GEMM_LDS + 64 * wc;
#pragma unroll
            for (int kk = 0; kk < GEMM_KC / 4; ++kk) {
                double fa[4], fb[8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) fa[mi] = As[kk * 4 * GEMM_LDS + 8 * mi];
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) fb[ni] = Bs[kk * 4 * GEMM_LDS + 8 * ni];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        // ---------------- epilogue ----------------
        const bool diag_tile = g.lower && (ti == tj);
        const double alpha = g.alpha;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cl = 64 * wc + 8 * ni + 2 * t + e;  // column inside the tile
                double* cp = g.C + (int64_t)(n0 + cl) * g.ldc + m0 + 32 * wr + gq;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int rl = 32 * wr + 8 * mi + gq;
                    if (diag_tile && rl < cl) continue;
                    double v = alpha * acc[mi][ni][e];
                    if (g.beta_one) v += cp[8 * mi];
                    cp[8 * mi] = v;
                }
            }
        }
    }
}

inline cudaError_t gemm_nt_prepare() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess) done = true;
    return e;
}

// algorithmic flops of one launch (what the roofline figure in bench.py is computed from)
inline double gemm_nt_flops(const GemmArgs& g) {
    const double tm = g.M / GEMM_BM, tn = g.N / GEMM_BN;
    if (g.k_from_tile) {  // U U^T on upper-triangular operands: tile (i, j<=i) contracts over K - 128 i
        double f = 0.0;
        for (int i = 0; i < (int)tm; ++i) f += (double)(i + 1) * (g.K - GEMM_BM * i);
        return 2.0 * GEMM_BM * GEMM_BN * f;
    }
    const double tiles = g.lower ? tm * (tm + 1) / 2 : tm * tn;
    return 2.0 * GEMM_BM * GEMM_BN * tiles * g.K;
}

// launches nothing when the problem is empty; returns the number of tiles launched
inline int64_t gemm_nt_launch(const GemmArgs& g, const LaunchCtx& ctx) {
    if (g.M <= 0 || g.N <= 0) return 0;
    const int64_t tm = g.M / GEMM_BM, tn = g.N / GEMM_BN;
    const int64_t tiles = g.lower ? tm * (tm + 1) / 2 : tm * tn;
    if (g.K <= 0) return 0;
    ProfScope ps(ctx, PROF_GEMM, gemm_nt_flops(g));
    gemm_nt_kernel<<<(unsigned)tiles, GEMM_THREADS, GEMM_SMEM_BYTES, ctx.st>>>(g);
    return tiles;
}

}  // namespace fgp
