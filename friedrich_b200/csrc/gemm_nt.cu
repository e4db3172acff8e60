// gemm_nt.cu — the fp64 tensor-pipe NT GEMM kernel (see gemm_nt.cuh for the contract and the execution model).
#include "gemm_nt.cuh"

// Experiment switches for tools/microbench/gemm_variants.cu (the production build leaves FGP_GEMM_EXP at 0):
//   1 = skip the epilogue (no C traffic)          2 = skip the operand loads and their barriers (main loop on stale smem)
//   3 = both                                      4 = epilogue without the C read (beta = 0)
#ifndef FGP_GEMM_EXP
#define FGP_GEMM_EXP 0
#endif

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <vector>

namespace fgp {

// One work item = one 64 x 128 half tile.
struct HalfTile {
    int m0, n0;      // origin in C
    int kbeg, nk;    // first contraction index, number of 16-column chunks
    int rbase;       // row of m0 inside its 128 x 128 tile (0 or 64)
    bool diag;       // lower mode, tile on the diagonal
};

__device__ __forceinline__ HalfTile half_tile(const GemmArgs& g, int item) {
    int ti, tj;
    gemm_tile_decode(g, item >> 1, ti, tj);
    HalfTile h;
    h.rbase = (item & 1) * GEMM_CTA_M;
    h.m0 = ti * GEMM_BM + h.rbase;
    h.n0 = tj * GEMM_BN;
    h.kbeg = g.k_from_tile ? GEMM_BM * max(ti, tj) : 0;
    h.nk = (g.K - h.kbeg) / GEMM_KC;
    h.diag = g.lower && (ti == tj);
    return h;
}

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_nt_kernel(const __grid_constant__ GemmArgs g, int n_items, const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GEMM_STAGES * GEMM_STAGE_DOUBLES * 8);
    uint64_t* empty = full + GEMM_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GEMM_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // ---------------- load cursor: runs GEMM_AHEAD chunks ahead of the compute cursor, across work items ----------------
    // two TMA tensor copies per k-chunk, issued by one lane: a (64+4) x 16 box of A and a (128+4) x 16 box of B. The 4 extra
    // rows are never read: they are the padding that makes the shared-memory column stride 68 / 132 doubles.
    int l_item = blockIdx.x;   // work item the next load belongs to
    int l_left = 0;            // chunks of that item not yet requested
    int l_m0 = 0, l_n0 = 0, l_k = 0;
    int l_idx = 0;             // running chunk index of the next load (stage = l_idx % STAGES)
    auto load_enter = [&](int item) {  // position the cursor on the first chunk of `item`
        const HalfTile h = half_tile(g, item);
        l_left = h.nk;
        l_m0 = h.m0;
        l_n0 = h.n0;
        l_k = h.kbeg;
        // pull the C half tile (128 columns x 512 B) into L2 now so the epilogue's reads do not pay DRAM latency
        if (g.beta_one) l2_prefetch(g.C + (int64_t)(h.n0 + lane + 32 * warp) * g.ldc + h.m0, GEMM_CTA_M * 8);
    };
    // every warp calls this once per chunk, in the same order; only `issuer` touches the barriers / TMA
    auto load_step = [&](bool issuer) {
        if (l_item >= n_items) return;
        if (issuer && !(FGP_GEMM_EXP & 2)) {
            const int s = l_idx % GEMM_STAGES;
            if (l_idx >= GEMM_STAGES) mbar_wait(&empty[s], ((l_idx / GEMM_STAGES) + 1) & 1);
            if (lane == 0) {
                double* dst = tiles + (size_t)s * GEMM_STAGE_DOUBLES;
                mbar_arrive_expect_tx(&full[s], GEMM_STAGE_DOUBLES * 8);
                tma_load_2d(dst, &tmA, l_m0, l_k, &full[s]);
                tma_load_2d(dst + GEMM_KC * GEMM_LDA, &tmB, l_n0, l_k, &full[s]);
            }
            __syncwarp();
        }
        ++l_idx;
        l_k += GEMM_KC;
        if (--l_left == 0) {
            l_item += gridDim.x;
            if (l_item < n_items) load_enter(l_item);
        }
    };
    if (l_item < n_items) load_enter(l_item);
#pragma unroll
    for (int a = 0; a < GEMM_AHEAD; ++a) load_step(warp == a);

    const int gq = lane >> 2, t = lane & 3;
    const int wr = warp & 1, wc = warp >> 1;
    int c_idx = 0;  // running chunk index of the compute cursor

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const HalfTile h = half_tile(g, item);
        // the top half of a diagonal tile has nothing on or below the diagonal in columns 64..127
        const bool idle = h.diag && h.rbase == 0 && wc == 1;
        double acc[4][8][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int it = 0; it < h.nk; ++it, ++c_idx) {
            // the warps take turns refilling the ring, GEMM_AHEAD chunks ahead of the one being consumed
            load_step((c_idx & (GEMM_WARPS - 1)) == warp);
            const int s = c_idx % GEMM_STAGES;
            if (!(FGP_GEMM_EXP & 2)) mbar_wait(&full[s], (c_idx / GEMM_STAGES) & 1);
            if (!idle) {
                const double* As = tiles + (size_t)s * GEMM_STAGE_DOUBLES + t * GEMM_LDA + 32 * wr + gq;
                const double* Bs = tiles + (size_t)s * GEMM_STAGE_DOUBLES + GEMM_KC * GEMM_LDA + t * GEMM_LDB + 64 * wc + gq;
#pragma unroll
                for (int kk = 0; kk < GEMM_KC / 4; ++kk) {
                    double fa[4], fb[8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) fa[mi] = As[kk * 4 * GEMM_LDA + 8 * mi];
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) fb[ni] = Bs[kk * 4 * GEMM_LDB + 8 * ni];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (idle) continue;
        if (FGP_GEMM_EXP & 1) {  // keep the accumulators observable without any C traffic
            double sum = 0.0;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
            if (sum == 12345.678) g.C[0] = sum;
            continue;
        }

        // ---------------- epilogue ----------------
        // Two 8-column groups at a time: all 16 C values of the batch are loaded before any is used, so the loads of a
        // batch are in flight together (the half tile was prefetched into L2 when its first chunk was requested).
        const double alpha = g.alpha;
        const int rbase = h.rbase + 32 * wr + gq;  // row inside the 128x128 tile of this lane's mi = 0 element
        double* cbase = g.C + (int64_t)h.n0 * g.ldc + h.m0 + 32 * wr + gq;
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
            double cv[2][2][4];
            if (g.beta_one && FGP_GEMM_EXP != 4) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cl = 64 * wc + 8 * (2 * nb2 + hh) + 2 * t + e;
                        const double* cp = cbase + (int64_t)cl * g.ldc;
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) cv[hh][e][mi] = __ldcg(cp + 8 * mi);
                    }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ni = 2 * nb2 + hh;
                    const int cl = 64 * wc + 8 * ni + 2 * t + e;
                    double* cp = cbase + (int64_t)cl * g.ldc;
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int rl = rbase + 8 * mi;
                        if (h.diag && rl < cl) continue;
                        double v = alpha * acc[mi][ni][e];
                        if (g.beta_one && FGP_GEMM_EXP != 4) v += cv[hh][e][mi];
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
    // two CTAs per SM need 2 x 100 KiB of shared memory: ask for the largest carve-out
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) done = true;
    return e;
}

int gemm_nt_occupancy() {
    int blocks = 0;
    if (gemm_nt_prepare() != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, gemm_nt_kernel, GEMM_THREADS, GEMM_SMEM_BYTES) != cudaSuccess)
        return -1;
    return blocks;
}

static int g_num_sms = 0;
static bool g_gemm_error = false;

bool gemm_nt_take_error() {
    const bool e = g_gemm_error;
    g_gemm_error = false;
    return e;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x cols column-major f64 operand (leading dimension ld), box = box_rows x GEMM_KC
static bool make_operand_map(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)GEMM_KC};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


void gemm_nt_plan(GemmArgs& g) {
    g.band_rows = GEMM_BAND_ROWS;
    g.n_bands = 1;
    g.band_prefix[0] = 0;
    g.band_prefix[1] = 0;
    if (!g.lower) return;
    const int tm = g.M / GEMM_BM, tn = g.N / GEMM_BN;
    const int PT = std::max(g.grp, 1), S = std::max(g.stride, 1);
    int R = GEMM_BAND_ROWS;
    while ((tm + R - 1) / R > GEMM_MAX_BANDS) R *= 2;
    const int nb = std::max(1, (tm + R - 1) / R);
    g.band_rows = R;
    g.n_bands = nb;
    std::vector<int64_t> cnt(nb, 0);
    for (int jl = 0; jl < tn; ++jl) {  // local column jl: tiles tj .. tm-1, spread over bands tj / R .. nb-1
        const int tj = (jl / PT) * S + jl % PT;
        if (tj >= tm) break;
        for (int r = tj / R; r < nb; ++r) {
            const int lo = std::max(r * R, tj), hi = std::min(tm, (r + 1) * R);
            cnt[r] += hi - lo;
        }
    }
    int64_t run = 0;
    for (int r = 0; r < nb; ++r) {
        g.band_prefix[r] = (int)run;
        run += cnt[r];
    }
    g.band_prefix[nb] = (int)run;
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
    GemmArgs p = g;
    gemm_nt_plan(p);
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    const int64_t items = 2 * tiles;  // 64 x 128 half tiles
    // one work item per CTA (the kernel also runs persistently with a smaller grid, but resident CTAs that never retire
    // would starve the look-ahead stream's panel kernels, and measured no faster: DESIGN.md "GEMM")
    const unsigned grid = (unsigned)items;
    // operand extents: rows beyond them arrive as zeros (only the padding rows of the last tile ever are). In lower mode the
    // rows of B are addressed by tile COLUMN positions of C, which reach M when the owned columns are strided (sharded).
    alignas(64) CUtensorMap tmA, tmB;
    if (!make_operand_map(&tmA, g.A, g.M, g.K, g.lda, GEMM_LDA) || !make_operand_map(&tmB, g.B, g.lower ? g.M : g.N, g.K, g.ldb, GEMM_LDB)) {
        if (!g_gemm_error) fprintf(stderr, "libfgp_sm100: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)\n", g.M, g.N, g.K);
        g_gemm_error = true;
        return 0;
    }
    ProfScope ps(ctx, PROF_GEMM, gemm_nt_flops(g));
    gemm_nt_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, ctx.st>>>(p, (int)items, tmA, tmB);
    return tiles;
}

}  // namespace fgp
