// gemm_nt.cu — the fp64 tensor-pipe NT GEMM kernel (see gemm_nt.cuh for the contract and the execution model).
#include "gemm_nt.cuh"

// Experiment switches for tools/microbench/gemm_variants.cu (the production build leaves FGP_GEMM_EXP at 0):
//   1 = skip the epilogue (no C traffic)          2 = skip the operand loads and their barriers (main loop on stale smem)
//   3 = both
#ifndef FGP_GEMM_EXP
#define FGP_GEMM_EXP 0
#endif

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <vector>

namespace fgp {

// Geometry of the two instantiations: CM = rows of the 128x128 tile one CTA computes.
//   CM = 64  two CTAs per tile, warps 2 x 2 of 32 x 64, 4-stage ring, 2 CTAs per SM: the throughput shape (big launches);
//   CM = 32  four CTAs per tile, warps 1 x 4 of 32 x 32, 3-stage ring, 3 CTAs per SM: launches of about one wave or less
//            (panel solves, in-panel and look-ahead updates, the solve steps of predict) — twice as many, half as long
//            CTAs spread a sub-wave launch evenly over the SMs (a 64-row launch of 150..296 CTAs leaves some SMs with two
//            CTAs and the rest with one, and lasts as long as the doubly loaded ones).
// Every output element sums its k-range in the same order in both, so the results are bit-identical.
template <int CM>
struct GemmShape {
    static constexpr int STAGES = (CM == 64) ? GEMM_STAGES : 3;
    static constexpr int LDA = CM + 4;                              // smem column stride of the A stage (doubles)
    static constexpr int STAGE_DOUBLES = GEMM_KC * (LDA + GEMM_LDB);
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8 + 2 * STAGES * 8;
    static constexpr int SUBS = GEMM_BM / CM;                       // CTAs per 128x128 tile
    static constexpr int WR = CM / 32;                              // warp grid: WR x WC
    static constexpr int WC = GEMM_WARPS / WR;
    static constexpr int NCOL = GEMM_BN / WC;                       // columns per warp
    static constexpr int NI = NCOL / 8;
    static constexpr int MIN_CTAS = (CM == 64) ? 2 : 3;
};

template <int CM>
__global__ void __launch_bounds__(GEMM_THREADS, GemmShape<CM>::MIN_CTAS)
gemm_nt_kernel(const __grid_constant__ GemmArgs g, const __grid_constant__ CUtensorMap tmA,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC) {
    using S = GemmShape<CM>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S::STAGES * S::STAGE_DOUBLES * 8);
    uint64_t* empty = full + S::STAGES;

    // work item = CM rows of a 128 x 128 tile: block group -> tile (ti, tj), block index inside the group -> row slice
    int ti, tj;
    gemm_tile_decode(g, blockIdx.x / S::SUBS, ti, tj);
    const int sub = blockIdx.x % S::SUBS;
    const int m0 = ti * GEMM_BM + sub * CM, n0 = tj * GEMM_BN;
    const int kbeg = g.k_from_tile ? GEMM_BM * (g.k_tile0 + max(ti, tj)) : 0;
    const int kend = g.k_upto_col ? min(g.K, GEMM_BN * (tj + 1)) : g.K;
    const int nk = (kend - kbeg) / GEMM_KC;
    const bool diag_tile = g.lower && (ti == tj);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (CM == 64 && g.stagger_ns > 0 && (int)blockIdx.x >= g.stagger_lo && (int)blockIdx.x < g.stagger_hi) {
        uint64_t t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            __nanosleep(2000);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < (uint64_t)g.stagger_ns);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GEMM_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Operand chunk j -> stage j % STAGES: two TMA tensor copies issued by one lane, a (CM+4) x 16 box of A and a (128+4) x 16
    // box of B. The 4 extra rows are never read: they are the padding that makes the shared-memory column stride CM+4 / 132.
    auto issue_load = [&](int j) {  // whole warp
        if (FGP_GEMM_EXP & 2) return;
        const int s = j % S::STAGES;
        if (j >= S::STAGES) mbar_wait(&empty[s], ((j / S::STAGES) + 1) & 1);  // chunk j - STAGES has left the stage
        if (lane == 0) {
            double* dst = tiles + (size_t)s * S::STAGE_DOUBLES;
            mbar_arrive_expect_tx(&full[s], S::STAGE_DOUBLES * 8);
            tma_load_2d(dst, &tmA, m0, kbeg + j * GEMM_KC, &full[s]);
            tma_load_2d(dst + GEMM_KC * S::LDA, &tmB, n0, kbeg + j * GEMM_KC, &full[s]);
        }
        __syncwarp();
    };
    if (warp < GEMM_AHEAD && warp < nk) issue_load(warp);
    // pull the C slice (128 columns x CM rows) towards L2 now: the epilogue's read-modify-write happens there
    if (g.beta_one) l2_prefetch(g.C + (int64_t)(n0 + lane + 32 * warp) * g.ldc + m0, CM * 8);

    const int gq = lane >> 2, t = lane & 3;
    const int wr = warp % S::WR, wc = warp / S::WR;
    // in a diagonal tile a warp whose columns all lie right of the slice's last row has nothing on or below the diagonal
    const bool idle = diag_tile && (S::NCOL * wc > sub * CM + CM - 1);
    double acc[4][S::NI][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < S::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    for (int it = 0; it < nk; ++it) {
        // the warps take turns refilling the ring, GEMM_AHEAD chunks ahead of the one being consumed
        if ((it & (GEMM_WARPS - 1)) == warp && it + GEMM_AHEAD < nk) issue_load(it + GEMM_AHEAD);
        const int s = it % S::STAGES;
        if (!(FGP_GEMM_EXP & 2)) mbar_wait(&full[s], (it / S::STAGES) & 1);
        if (!idle) {
            const double* As = tiles + (size_t)s * S::STAGE_DOUBLES + t * S::LDA + 32 * wr + gq;
            const double* Bs = tiles + (size_t)s * S::STAGE_DOUBLES + GEMM_KC * S::LDA + t * GEMM_LDB + S::NCOL * wc + gq;
#pragma unroll
            for (int kk = 0; kk < GEMM_KC / 4; ++kk) {
                double fa[4], fb[S::NI];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) fa[mi] = As[kk * 4 * S::LDA + 8 * mi];
#pragma unroll
                for (int ni = 0; ni < S::NI; ++ni) fb[ni] = Bs[kk * 4 * GEMM_LDB + 8 * ni];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < S::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (FGP_GEMM_EXP & 1) {  // keep the accumulators observable without any C traffic
        double sum = 0.0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < S::NI; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
        if (sum == 12345.678) g.C[0] = sum;
        return;
    }

    // ---------------- epilogue ----------------
    // alpha * acc is staged in the (now idle) operand ring as a dense CM x 128 column-major image and handed to the TMA
    // engine in ONE tensor operation: a reduce-add into C (beta = 1: the f64 addition happens at the L2, the SM never reads
    // C) or a plain store (beta = 0). Elements above the diagonal of a diagonal tile are staged as zeros (C + 0 = C).
    __syncthreads();  // every warp has consumed every chunk: no TMA write into the ring is pending, nobody reads it any more
    {
        const double alpha = g.alpha;
        double* stage = tiles + 32 * wr + gq;
        const int rbase = sub * CM + 32 * wr + gq;  // row inside the 128x128 tile of this lane's mi = 0 element
#pragma unroll
        for (int ni = 0; ni < S::NI; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cl = S::NCOL * wc + 8 * ni + 2 * t + e;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const bool above = diag_tile && (rbase + 8 * mi < cl);
                    stage[cl * CM + 8 * mi] = (idle || above) ? 0.0 : alpha * acc[mi][ni][e];
                }
            }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (g.beta_one) tma_reduce_add_2d(&tmC, m0, n0, tiles);
        else tma_store_2d(&tmC, m0, n0, tiles);
        tma_commit_group();
        tma_wait_group_read0();  // the staging image must outlive the TMA read
    }
}

template <int CM>
static cudaError_t prepare_shape() {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_nt_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmShape<CM>::SMEM_BYTES);
    // several CTAs per SM need most of the SM's shared memory: ask for the largest carve-out
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<CM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return e;
}

cudaError_t gemm_nt_prepare() {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (done) return cudaSuccess;
    cudaError_t e = prepare_shape<64>();
    if (e == cudaSuccess) e = prepare_shape<32>();
    if (e == cudaSuccess) done = true;
    return e;
}

int gemm_nt_occupancy(int cta_rows) {
    int blocks = 0;
    if (gemm_nt_prepare() != cudaSuccess) return -1;
    const cudaError_t e =
        (cta_rows == 32)
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, gemm_nt_kernel<32>, GEMM_THREADS, GemmShape<32>::SMEM_BYTES)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, gemm_nt_kernel<64>, GEMM_THREADS, GemmShape<64>::SMEM_BYTES);
    return e == cudaSuccess ? blocks : -1;
}

// Per-device state (a process may hold models on several GPUs and drive them from several threads): the SM count that
// picks the CTA shape, and the "a launch could not be set up" flag that end_timed() of a model on THAT device consumes.
static std::atomic<int> g_num_sms[64];
static std::atomic<bool> g_gemm_error[64];

static int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64) ? dev : 0;
}

void gemm_nt_flag_error() { g_gemm_error[current_device_slot()].store(true); }

bool gemm_nt_take_error() { return g_gemm_error[current_device_slot()].exchange(false); }

int gemm_nt_num_sms() {
    const int slot = current_device_slot();
    int n = g_num_sms[slot].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot);
        if (n <= 0) n = 148;
        g_num_sms[slot].store(n, std::memory_order_relaxed);
    }
    return n;
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

bool make_tile_map(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
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
    const int skip = std::min(std::max(g.row_skip, 0), tm);
    int R = GEMM_BAND_ROWS;
    while ((tm + R - 1) / R > GEMM_MAX_BANDS) R *= 2;
    const int nb = std::max(1, (tm - skip + R - 1) / R);
    g.band_rows = R;
    g.n_bands = nb;
    std::vector<int64_t> cnt(nb, 0);
    for (int jl = 0; jl < tn; ++jl) {  // local column jl: tiles max(tj, skip) .. tm-1, spread over the bands from its first row on
        const int tj = (jl / PT) * S + jl % PT;
        if (tj >= tm) break;
        const int first = std::max(tj, skip);
        for (int r = (first - skip) / R; r < nb; ++r) {
            const int lo = std::max(skip + r * R, first), hi = std::min(tm, skip + (r + 1) * R);
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
        const int64_t tj = (jl / PT) * S + jl % PT, first = std::max<int64_t>(tj, g.row_skip);
        if (tm - first > 0) tiles += tm - first;
    }
    return tiles;
}

double gemm_nt_flops(const GemmArgs& g) {
    if (g.k_from_tile) {  // U U^T on upper-triangular operands: tile (i, j<=i) contracts over K - 128 (k_tile0 + i)
        GemmArgs p = g;
        gemm_nt_plan(p);
        const int64_t tiles = gemm_nt_tiles(g);
        double f = 0.0;
        for (int64_t b = 0; b < tiles; ++b) {
            int ti, tj;
            gemm_tile_decode(p, (int)b, ti, tj);
            f += (double)(g.K - GEMM_BM * (g.k_tile0 + std::max(ti, tj)));
        }
        return 2.0 * GEMM_BM * GEMM_BN * f;
    }
    if (g.k_upto_col && !g.lower) {  // A W^T with W lower block-triangular: tile column j contracts over min(K, 128 (j + 1))
        double f = 0.0;
        for (int j = 0; j < g.N / GEMM_BN; ++j) f += std::min<double>(g.K, GEMM_BN * (j + 1));
        return 2.0 * GEMM_BM * GEMM_BN * (g.M / GEMM_BM) * f;
    }
    return 2.0 * GEMM_BM * GEMM_BN * (double)gemm_nt_tiles(g) * g.K;
}

int gemm_nt_cta_rows(int64_t tiles, int num_sms) {
    const int64_t i64 = 2 * tiles, S = num_sms;
    return ((2 * i64 <= S) || (i64 > S && 2 * i64 <= 3 * S)) ? 32 : 64;
}

int64_t gemm_nt_launch(const GemmArgs& g, const LaunchCtx& ctx) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const int64_t tiles = gemm_nt_tiles(g);
    if (tiles <= 0) return 0;
    GemmArgs p = g;
    gemm_nt_plan(p);
    const int num_sms = gemm_nt_num_sms();
    // shape: 64-row CTAs for throughput; 32-row CTAs where halving the CTAs lowers the heaviest SM's load (see GemmShape).
    // With i = 64-row items and S SMs the heaviest SM carries ceil(i / S) units, with 32-row CTAs ceil(2 i / S) / 2: that is
    // less for i <= S/2 (0.5 vs 1) and for S < i <= 1.5 S (1.5 vs 2); elsewhere the 64-row shape is as balanced and cheaper
    // (measured: M = 4096, N = 128, K = 384: 45.7 -> 25.2 us; M = 16384, same N, K: 64 us with 64 rows, 74 us with 32).
#ifdef FGP_GEMM_FORCE_CM
    const bool small = (FGP_GEMM_FORCE_CM == 32);
#else
    const bool small = gemm_nt_cta_rows(tiles, num_sms) == 32;
#endif
    const int cm = small ? 32 : 64;
    const int64_t items = tiles * (GEMM_BM / cm);
    // one work item per CTA (a persistent variant measured no faster and its never-retiring CTAs starve the look-ahead
    // stream's panel kernels: DESIGN.md "GEMM")
    const unsigned grid = (unsigned)items;
    // de-synchronise the two resident CTAs of every SM (they would otherwise start, and therefore finish, together for the
    // whole launch) by half the lifetime of a CTA pair: 2 x 64x128xK x 2 flop at 128 flop/clk/SM = K x 131 ns, plus overheads
#ifndef FGP_GEMM_NO_STAGGER
    if (!small && items >= 8 * (int64_t)num_sms) {
        p.stagger_lo = num_sms;
        p.stagger_hi = 2 * num_sms;
        p.stagger_ns = (g.K - (g.k_from_tile ? g.K / 2 : 0)) * 70;
    }
#endif
    // operand extents: rows beyond them arrive as zeros (only the padding rows of the last tile ever are). In lower mode the
    // rows of B are addressed by tile COLUMN positions of C, which reach M when the owned columns are strided (sharded).
    alignas(64) CUtensorMap tmA, tmB, tmC;
    const int64_t ncols = g.lower ? g.M : g.N;
    if (!make_tile_map(&tmA, g.A, g.M, g.K, g.lda, cm + 4, GEMM_KC) || !make_tile_map(&tmB, g.B, ncols, g.K, g.ldb, GEMM_LDB, GEMM_KC) ||
        !make_tile_map(&tmC, g.C, g.M, ncols, g.ldc, cm, GEMM_BN)) {
        if (!g_gemm_error[current_device_slot()].exchange(true))
            fprintf(stderr, "libfgp_sm100: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)\n", g.M, g.N, g.K);
        return 0;
    }
    ProfScope ps(ctx, PROF_GEMM, gemm_nt_flops(g));
    if (small) gemm_nt_kernel<32><<<grid, GEMM_THREADS, GemmShape<32>::SMEM_BYTES, ctx.st>>>(p, tmA, tmB, tmC);
    else gemm_nt_kernel<64><<<grid, GEMM_THREADS, GemmShape<64>::SMEM_BYTES, ctx.st>>>(p, tmA, tmB, tmC);
    return tiles;
}

}  // namespace fgp
