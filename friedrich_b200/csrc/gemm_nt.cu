// gemm_nt.cu — the fp64 tensor-pipe NT GEMM kernel (see gemm_nt.cuh for the contract and the execution model).
#include "gemm_nt.cuh"

#include <algorithm>

namespace fgp {

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
            lower_tile_decode(tm, max(g.grp, 1), max(g.stride, 1), b, ti, tj);
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
        if (g.beta_one) {
            // pull the C tile (128 columns x 1 KiB) into L2 now so the epilogue's reads do not pay DRAM latency
#pragma unroll
            for (int c = 0; c < 4; ++c)
                l2_prefetch(g.C + (int64_t)(n0 + lane + 32 * c) * g.ldc + m0, GEMM_BM * 8);
        }
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
        // Two 8-column groups at a time: all 16 C values of the batch are loaded before any is used, so the loads of a
        // batch are in flight together (the tile was prefetched into L2 by the producer at kernel start).
        const bool diag_tile = g.lower && (ti == tj);
        const double alpha = g.alpha;
        double* cbase = g.C + (int64_t)n0 * g.ldc + m0 + 32 * wr + gq;
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
            double cv[2][2][4];
            if (g.beta_one) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cl = 64 * wc + 8 * (2 * nb2 + h) + 2 * t + e;
                        const double* cp = cbase + (int64_t)cl * g.ldc;
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) cv[h][e][mi] = __ldcg(cp + 8 * mi);
                    }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ni = 2 * nb2 + h;
                    const int cl = 64 * wc + 8 * ni + 2 * t + e;
                    double* cp = cbase + (int64_t)cl * g.ldc;
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int rl = 32 * wr + 8 * mi + gq;
                        if (diag_tile && rl < cl) continue;
                        double v = alpha * acc[mi][ni][e];
                        if (g.beta_one) v += cv[h][e][mi];
                        __stcg(cp + 8 * mi, v);
                    }
                }
        }
    }
}

cudaError_t gemm_nt_prepare() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess) done = true;
    return e;
}

int64_t gemm_nt_tiles(const GemmArgs& g) {
    const int64_t tm = g.M / GEMM_BM, tn = g.N / GEMM_BN;
    if (!g.lower) return tm * tn;
    const int64_t PT = std::max(g.grp, 1), S = std::max(g.stride, 1);
    int64_t tiles = 0;
    for (int64_t jl = 0; jl < tn; ++jl) {  // local tile column jl sits at tile column (jl / PT) * S + jl % PT
        const int64_t tj = (jl / PT) * S + jl % PT;
        if (tm - tj > 0) tiles += tm - tj;
    }
    return tiles;
}

double gemm_nt_flops(const GemmArgs& g) {
    if (g.k_from_tile) {  // U U^T on upper-triangular operands: tile (i, j<=i) contracts over K - 128 i
        const int tm = g.M / GEMM_BM;
        double f = 0.0;
        for (int i = 0; i < tm; ++i) f += (double)(i + 1) * (g.K - GEMM_BM * i);
        return 2.0 * GEMM_BM * GEMM_BN * f;
    }
    return 2.0 * GEMM_BM * GEMM_BN * (double)gemm_nt_tiles(g) * g.K;
}

int64_t gemm_nt_launch(const GemmArgs& g, const LaunchCtx& ctx) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const int64_t tiles = gemm_nt_tiles(g);
    if (tiles <= 0) return 0;
    ProfScope ps(ctx, PROF_GEMM, gemm_nt_flops(g));
    gemm_nt_kernel<<<(unsigned)tiles, GEMM_THREADS, GEMM_SMEM_BYTES, ctx.st>>>(g);
    return tiles;
}

}  // namespace fgp
