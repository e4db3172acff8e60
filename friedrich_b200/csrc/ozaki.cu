// ozaki.cu — exact-integer trailing update on tcgen05 / TMEM (contract, arithmetic and data layout: ozaki.cuh).
#include "ozaki.cuh"

#include <algorithm>
#include <cstdio>

namespace fgp {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// slicing: one CTA per row tile (128 rows), 512 threads = 4 k-quarters x 128 rows
__global__ void __launch_bounds__(512) ozaki_slice_kernel(const double* __restrict__ P, int64_t ld, int K, int8_t* __restrict__ digits,
                                                          double* __restrict__ scale) {
    __shared__ double red[4][128];
    __shared__ int bad[128];
    const int r = threadIdx.x & 127, q = threadIdx.x >> 7;
    const int64_t T = blockIdx.x;
    const int kq = K >> 2, k0 = q * kq, KS = K / OZ_KSTEP;
    const double* p = P + T * 128 + r + (int64_t)k0 * ld;
    if (q == 0) bad[r] = 0;
    __syncthreads();
    double m = 0.0;
    bool nonfinite = false;
    for (int k = 0; k < kq; ++k) {
        const double a = fabs(p[(int64_t)k * ld]);
        nonfinite |= !(a <= 1.7976931348623157e308);
        m = fmax(m, a);
    }
    red[q][r] = m;
    if (nonfinite) bad[r] = 1;
    __syncthreads();
    m = fmax(fmax(red[0][r], red[1][r]), fmax(red[2][r], red[3][r]));
    // max |p| < 2^e
    int e = ((__double2hiint(m) >> 20) & 0x7ff) - 1022;
    const bool zero = (m == 0.0) || e < -900;
    const bool nan = bad[r] || e > 900;
    if (q == 0) scale[T * 128 + r] = nan ? __longlong_as_double(0x7ff8000000000000ll) : zero ? 0.0 : __hiloint2double((1023 + e - 30) << 20, 0);
    const double up = (zero || nan) ? 0.0 : __hiloint2double((1023 + 55 - e) << 20, 0);  // 2^(55 - e)
    for (int s = k0 / OZ_KSTEP; s < (k0 + kq) / OZ_KSTEP; ++s) {
#pragma unroll 1
        for (int kh = 0; kh < 2; ++kh) {
            uint32_t w[OZ_SLICES][4];
#pragma unroll
            for (int i = 0; i < OZ_SLICES; ++i) w[i][0] = w[i][1] = w[i][2] = w[i][3] = 0u;
            const double* src = P + T * 128 + r + (int64_t)(s * OZ_KSTEP + kh * 16) * ld;
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {
                long long X = __double2ll_rn(src[(int64_t)kb * ld] * up);
#pragma unroll
                for (int i = OZ_SLICES - 1; i >= 1; --i) {
                    const int d = (((int)X & 127) ^ 64) - 64;   // balanced digit in [-64, 63]
                    X = (X - d) >> 7;
                    w[i][kb >> 2] |= (uint32_t)(d & 0xff) << (8 * (kb & 3));
                }
                w[0][kb >> 2] |= (uint32_t)((int)X & 0xff) << (8 * (kb & 3));   // |top digit| <= 65
            }
            int8_t* dst = digits + ((T * KS + s) * OZ_SLICES) * (int64_t)OZ_BLOCK_BYTES + kh * 2048 + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
            for (int i = 0; i < OZ_SLICES; ++i)
                *reinterpret_cast<uint4*>(dst + (int64_t)i * OZ_BLOCK_BYTES) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct OzArgs {
    const int8_t* SA;
    const int8_t* SB;
    const double* scA;
    const double* scB;
    int KS;          // k-steps (K / 32)
    int tiles, tpc;  // tiles of the launch, tiles per CTA (CTA b: tiles [b * tpc, (b + 1) * tpc))
    uint32_t lbo, sbo;
};

// mbarrier wait with a watchdog: a protocol error traps instead of hanging the device
__device__ __forceinline__ void oz_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) {
            printf("ozaki_update_kernel: barrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, a, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, no swizzle: start address, leading-dimension (k direction) and stride (8-row group) byte offsets in 16 B units
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

// kind::i8, D = S32, A = B = signed 8 bit, both K-major, N = 128, M = 128
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_update_kernel(const __grid_constant__ GemmArgs g, const __grid_constant__ CUtensorMap tmC, const OzArgs o) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    double* staging = reinterpret_cast<double*>(smem + OZ_STAGES * OZ_STAGE_BYTES);
    unsigned char* tail = smem + OZ_STAGES * OZ_STAGE_BYTES + OZ_STAGING_BYTES;
    double* colsc = reinterpret_cast<double*>(tail);                    // 128 doubles
    uint64_t* full = reinterpret_cast<uint64_t*>(tail + 1024);          // [OZ_STAGES]
    uint64_t* empty = full + OZ_STAGES;                                 // [OZ_STAGES]
    uint64_t* tmem_full = empty + OZ_STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = blockIdx.x * o.tpc, t_end = min(t_begin + o.tpc, o.tiles);

    if (threadIdx.x == 32) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== producer: bulk copies of the digit blocks =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                int ti, tj;
                gemm_tile_decode(g, t, ti, tj);
                const int8_t* a0 = o.SA + (int64_t)ti * o.KS * OZ_PART_BYTES;
                const int8_t* b0 = o.SB + (int64_t)tj * o.KS * OZ_PART_BYTES;
                for (int pass = 0; pass < 2; ++pass) {
                    const uint32_t bytes = pass == 0 ? OZ_PART_BYTES / 2 : OZ_PART_BYTES;
                    for (int ks = 0; ks < o.KS; ++ks) {
                        oz_wait(&empty[stage], phase ^ 1);
                        unsigned char* dst = ring + stage * OZ_STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                        tma_load_1d(dst, a0 + (int64_t)ks * OZ_PART_BYTES, bytes, &full[stage]);
                        tma_load_1d(dst + OZ_PART_BYTES, b0 + (int64_t)ks * OZ_PART_BYTES, bytes, &full[stage]);
                        if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, npass = 0;
            for (int t = t_begin; t < t_end; ++t) {
                for (int pass = 0; pass < 2; ++pass, ++npass) {
                    oz_wait(tmem_empty, (npass & 1) ^ 1);   // the epilogue has drained the four accumulators
                    tc_fence_after();
                    const int g0 = pass * 4;
                    for (int ks = 0; ks < o.KS; ++ks) {
                        oz_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + stage * OZ_STAGE_BYTES), sb = sa + OZ_PART_BYTES;
                        for (int gg = 0; gg < 4; ++gg) {
                            const int grp = g0 + gg;
                            const uint32_t d = tmem_base + gg * 128;
                            for (int i = 0; i <= grp; ++i) {
                                const int j = grp - i;
                                umma_i8(d, oz_desc(sa + i * OZ_BLOCK_BYTES, o.lbo, o.sbo), oz_desc(sb + j * OZ_BLOCK_BYTES, o.lbo, o.sbo),
                                        OZ_IDESC, (ks > 0 || i > 0) ? 1u : 0u);
                            }
                        }
                        tc_commit(&empty[stage]);   // the stage is free once these MMAs have read it
                        if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tmem_full);           // accumulators complete
                }
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const int etid = (warp - 2) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t npass = 0, nchunk = 0;
        for (int t = t_begin; t < t_end; ++t) {
            int ti, tj;
            gemm_tile_decode(g, t, ti, tj);
            const bool diag = g.lower && ti == tj;
            const int m0 = ti * 128, n0 = tj * 128;
            colsc[etid] = o.scB[n0 + etid];
            const double rs = o.scA[m0 + row];
            epi_bar();
            for (int pass = 0; pass < 2; ++pass, ++npass) {
                // pass 0 carries the groups 0..3: weight 128^4 = 2^28 over pass 1; alpha = -1; sc = 2^(e-30): 2^(e_r+e_c-61) = rs*cs/2
                const double rf = rs * (pass == 0 ? -134217728.0 : -0.5);
                oz_wait(tmem_full, npass & 1);
                tc_fence_after();
                for (int ch = 0; ch < 128 / OZ_STAGING_COLS; ++ch, ++nchunk) {
                    int v[4][16];
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) tmem_ld16(lane_addr + gg * 128 + ch * OZ_STAGING_COLS, v[gg]);
                    tmem_ld_wait();
                    if (ch == 128 / OZ_STAGING_COLS - 1) {   // TMEM is free for the next pass as soon as the last load has landed
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty);
                    }
                    double* st = staging + (nchunk & 1) * (128 * OZ_STAGING_COLS) + row;
#pragma unroll
                    for (int c = 0; c < OZ_STAGING_COLS; ++c) {
                        const long long s = (long long)v[0][c] * 2097152ll + (long long)v[1][c] * 16384ll + (long long)v[2][c] * 128ll + (long long)v[3][c];
                        const int cl = ch * OZ_STAGING_COLS + c;
                        const double val = (double)s * (rf * colsc[cl]);
                        st[c * 128] = (diag && row < cl) ? 0.0 : val;
                    }
                    fence_proxy_async_smem();
                    if (etid == 0) {
                        // every earlier reduce has read its staging image; at the first chunk of a pass also: has been PERFORMED,
                        // so that the two passes' additions to an element of C are applied in a fixed order
                        if (ch == 0) tma_wait_group0();
                        else tma_wait_group_read0();
                    }
                    epi_bar();
                    if (etid == 0) {
                        tma_reduce_add_2d(&tmC, m0, n0 + ch * OZ_STAGING_COLS, staging + (nchunk & 1) * (128 * OZ_STAGING_COLS));
                        tma_commit_group();
                    }
                }
            }
        }
        if (etid == 0) tma_wait_group0();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

cudaError_t ozaki_prepare() {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(ozaki_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
    if (e == cudaSuccess) done = true;
    return e;
}

void ozaki_slice_launch(const double* P, int64_t ld, int64_t rows, int K, int8_t* digits, double* scale, const LaunchCtx& ctx) {
    if (rows <= 0 || K <= 0) return;
    ProfScope ps(ctx, PROF_OTHER, 0.0);
    ozaki_slice_kernel<<<(unsigned)(rows / 128), 512, 0, ctx.st>>>(P, ld, K, digits, scale);
}

int64_t ozaki_update_launch(const GemmArgs& g, const int8_t* digitsA, const double* scaleA, const int8_t* digitsB,
                            const double* scaleB, int tiles_per_cta, const LaunchCtx& ctx, uint32_t lbo, uint32_t sbo) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const int64_t tiles = gemm_nt_tiles(g);
    if (tiles <= 0) return 0;
    if (ozaki_prepare() != cudaSuccess) {
        gemm_nt_flag_error();
        return 0;
    }
    GemmArgs p = g;
    gemm_nt_plan(p);
    alignas(64) CUtensorMap tmC;
    const int64_t ncols = g.lower ? g.M : g.N;
    if (!make_tile_map(&tmC, g.C, g.M, ncols, g.ldc, 128, OZ_STAGING_COLS)) {
        gemm_nt_flag_error();
        return 0;
    }
    OzArgs o{};
    o.SA = digitsA; o.SB = digitsB; o.scA = scaleA; o.scB = scaleB;
    o.KS = g.K / OZ_KSTEP;
    o.tiles = (int)tiles;
    const int num_sms = gemm_nt_num_sms();
    int tpc = tiles_per_cta;
    if (tpc <= 0) tpc = (int)std::max<int64_t>(1, std::min<int64_t>(4, tiles / (2 * num_sms)));
    o.tpc = tpc;
    o.lbo = lbo; o.sbo = sbo;
    const unsigned grid = (unsigned)((tiles + tpc - 1) / tpc);
    ProfScope ps(ctx, PROF_GEMM, gemm_nt_flops(g));
    ozaki_update_kernel<<<grid, OZ_THREADS, OZ_SMEM_BYTES, ctx.st>>>(p, tmC, o);
    return tiles;
}

}  // namespace fgp
