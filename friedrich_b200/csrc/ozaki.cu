// ozaki.cu — exact-integer trailing update on tcgen05 / TMEM (contract, arithmetic and data layout: ozaki.cuh).
#include "ozaki.cuh"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace fgp {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// slicing, step 1: row scales.  One CTA per row tile (128 rows), 1024 threads = 8 k-groups x 128 rows (coalesced along rows).
__global__ void __launch_bounds__(1024) ozaki_rowmax_kernel(const double* __restrict__ P, int64_t ld, int K, double* __restrict__ scale) {
    __shared__ double red[8][128];
    __shared__ int bad[128];
    const int r = threadIdx.x & 127, q = threadIdx.x >> 7;
    const int64_t T = blockIdx.x;
    const int kq = K >> 3;
    const double* p = P + T * 128 + r + (int64_t)(q * kq) * ld;
    if (q == 0) bad[r] = 0;
    __syncthreads();
    double m = 0.0;
    bool nonfinite = false;
#pragma unroll 8
    for (int k = 0; k < kq; ++k) {
        const double a = fabs(p[(int64_t)k * ld]);
        nonfinite |= !(a <= 1.7976931348623157e308);
        m = fmax(m, a);
    }
    red[q][r] = m;
    if (nonfinite) bad[r] = 1;
    __syncthreads();
    if (q == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmax(m, red[i][r]);
        const int e = ((__double2hiint(m) >> 20) & 0x7ff) - 1022;   // max |p| < 2^e
        const bool zero = (m == 0.0) || e < -900;
        const bool nan = bad[r] || e > 900;
        scale[T * 128 + r] = nan ? __longlong_as_double(0x7ff8000000000000ll) : zero ? 0.0 : __hiloint2double((1023 + e - 30) << 20, 0);
    }
}

// slicing, step 2: digits.  CTA (T, s) = row tile T, k-step s; 256 threads = 2 k-halves x 128 rows; every thread turns 16
// consecutive k of its row into 8 x 16 digit bytes (one 16-byte store per slice, coalesced across the rows of a warp).
__global__ void __launch_bounds__(256) ozaki_digits_kernel(const double* __restrict__ P, int64_t ld, int KS, const double* __restrict__ scale,
                                                           int8_t* __restrict__ digits) {
    const int r = threadIdx.x & 127, kh = threadIdx.x >> 7;
    const int64_t T = blockIdx.x;
    const int s = blockIdx.y;
    const double sc = scale[T * 128 + r];
    // scale = 2^(e - 30): digits are taken from rint(p * 2^(55 - e)) = rint(p * 2^25 / scale)
    const double up = (sc > 0.0) ? __hiloint2double(((2046 + 25) << 20) - __double2hiint(sc), 0) : 0.0;   // hi word of 2^x = (1023 + x) << 20
    const double* src = P + T * 128 + r + (int64_t)(s * OZ_KSTEP + kh * 16) * ld;
    double a[16];
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) a[kb] = src[(int64_t)kb * ld];
    uint32_t w[OZ_SLICES][4];
#pragma unroll
    for (int i = 0; i < OZ_SLICES; ++i) w[i][0] = w[i][1] = w[i][2] = w[i][3] = 0u;
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) {
        long long X = __double2ll_rn(a[kb] * up);
#pragma unroll
        for (int i = OZ_SLICES - 1; i >= 1; --i) {
            const int d = (((int)X & 127) ^ 64) - 64;   // balanced digit in [-64, 63]
            X = (X - d) >> 7;
            w[i][kb >> 2] |= (uint32_t)(d & 0xff) << (8 * (kb & 3));
        }
        w[0][kb >> 2] |= (uint32_t)((int)X & 0xff) << (8 * (kb & 3));   // |top digit| <= 65
    }
    int8_t* dst = digits + ((T * KS + s) * OZ_SLICES) * (int64_t)OZ_BLOCK_BYTES + r * 32 + ((kh ^ ((r >> 2) & 1)) << 4);
#pragma unroll
    for (int i = 0; i < OZ_SLICES; ++i)
        *reinterpret_cast<uint4*>(dst + (int64_t)i * OZ_BLOCK_BYTES) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
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
    double alpha;    // C += alpha * P Q^T (the factorisation uses -1; K^-1 = U U^T uses +1 on a zeroed C)
    int exp;         // measurement switches (tools/ozaki_check.py): 1 = no epilogue arithmetic / stores, 2 = no operand copies, 4 = no MMAs
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
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tma_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// kind::i8, D = S32, A = B = signed 8 bit, both K-major, N = 128, M = 128
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// the MMAs of one k-step of pass PASS (groups 4 PASS .. 4 PASS + 3), fully unrolled: the descriptors differ from the stage's
// base descriptors by compile-time constants (slice i sits i * 4096 bytes = i * 256 descriptor units into its part)
template <int PASS>
__device__ __forceinline__ void issue_kstep(uint32_t tmem_base, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t not_first) {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
        constexpr int G0 = PASS * 4;
#pragma unroll
        for (int i = 0; i <= G0 + gg; ++i) {
            const int j = G0 + gg - i;
            umma_i8(tmem_base + gg * 128, desc64(a_lo + i * (OZ_BLOCK_BYTES >> 4), hi), desc64(b_lo + j * (OZ_BLOCK_BYTES >> 4), hi), OZ_IDESC,
                    i > 0 ? 1u : not_first);
        }
    }
}
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_update_kernel(const __grid_constant__ GemmArgs g, const __grid_constant__ CUtensorMap tmC, const OzArgs o) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    unsigned char* tail = smem + OZ_STAGES * OZ_STAGE_BYTES + OZ_STAGING_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);                 // [OZ_STAGES]
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
        mbar_init(tmem_empty, OZ_EPI_WARPS);
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
        // ===== producer: bulk copies of the digit blocks (the whole warp walks the loop, one elected lane issues) =====
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            int ti, tj;
            gemm_tile_decode(g, t, ti, tj);
            const int8_t* a0 = o.SA + (int64_t)ti * o.KS * OZ_PART_BYTES;
            const int8_t* b0 = o.SB + (int64_t)tj * o.KS * OZ_PART_BYTES;
            // k_from_tile (U U^T on upper-triangular operands): the contraction starts at k = 128 (k_tile0 + max(ti, tj))
            const int ks0 = g.k_from_tile ? (128 / OZ_KSTEP) * (g.k_tile0 + max(ti, tj)) : 0;
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t bytes = pass == 0 ? OZ_PART_BYTES / 2 : OZ_PART_BYTES;
                for (int ks = ks0; ks < o.KS; ++ks) {
                    oz_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        unsigned char* dst = ring + stage * OZ_STAGE_BYTES;
                        if (o.exp & 2) {
                            mbar_arrive(&full[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                            tma_load_1d(dst, a0 + (int64_t)ks * OZ_PART_BYTES, bytes, &full[stage]);
                            tma_load_1d(dst + OZ_PART_BYTES, b0 + (int64_t)ks * OZ_PART_BYTES, bytes, &full[stage]);
                        }
                    }
                    __syncwarp();
                    if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (uniform control flow and descriptor arithmetic), one elected lane
        // issues the tcgen05.mma / tcgen05.commit instructions =====
        int stage = 0;
        uint32_t phase = 0, npass = 0;
        const uint32_t hi = (uint32_t)(oz_desc(0, o.lbo, o.sbo) >> 32), lbo_field = (o.lbo >> 4) << 16;
        const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | lbo_field;
        for (int t = t_begin; t < t_end; ++t) {
            int ks0 = 0;
            if (g.k_from_tile) {
                int ti, tj;
                gemm_tile_decode(g, t, ti, tj);
                ks0 = (128 / OZ_KSTEP) * (g.k_tile0 + max(ti, tj));
            }
            for (int pass = 0; pass < 2; ++pass, ++npass) {
                oz_wait(tmem_empty, (npass & 1) ^ 1);   // the epilogue has drained the four accumulators
                tc_fence_after();
                for (int ks = ks0; ks < o.KS; ++ks) {
                    oz_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo = ring_lo + stage * (OZ_STAGE_BYTES >> 4), b_lo = a_lo + (OZ_PART_BYTES >> 4);
                        if (!(o.exp & 4)) {
                            if (pass == 0) issue_kstep<0>(tmem_base, a_lo, b_lo, hi, ks > ks0);
                            else issue_kstep<1>(tmem_base, a_lo, b_lo, hi, ks > ks0);
                        }
                        tc_commit(&empty[stage]);                    // the stage is free once these MMAs have read it
                        if (ks == o.KS - 1) tc_commit(tmem_full);    // accumulators complete
                    }
                    __syncwarp();
                    if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4 (32 tile rows), column half = (warp - 2) / 4 =====
        const int quarter = warp & 3, row = quarter * 32 + lane, half = (warp - 2) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        double* stg = reinterpret_cast<double*>(smem + OZ_STAGES * OZ_STAGE_BYTES) + (warp - 2) * (32 * 16);   // this warp's 4 KB image
        uint32_t npass = 0;
        for (int t = t_begin; t < t_end; ++t) {
            int ti, tj;
            gemm_tile_decode(g, t, ti, tj);
            const bool diag = g.lower && ti == tj;
            const int m0 = ti * 128, n0 = tj * 128;
            const double rs = o.scA[m0 + row];
            for (int pass = 0; pass < 2; ++pass, ++npass) {
                // pass 0 carries the groups 0..3: weight 128^4 = 2^28 over pass 1; sc = 2^(e-30): 2^(e_r+e_c-61) = rs*cs/2
                const double rf = rs * o.alpha * (pass == 0 ? 134217728.0 : 0.5);
                oz_wait(tmem_full, npass & 1);
                tc_fence_after();
                // phase 1 (TMEM busy, the MMA warp waits): the warp's 32 rows x 64 columns of the four accumulators -> one exact f64
                // per element in registers.  Conversions to f64 run at 16 / clk / SM, so the integers are combined pairwise in 64-bit
                // integer arithmetic and turned into doubles by the 2^52 trick (exact for |u| < 2^51): 3 f64 operations per element
                double x[64];
                int v[2][4][4];
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) tmem_ld4(lane_addr + gg * 128 + half * 64, v[0][gg]);
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {   // the next 4 columns' TMEM loads are in flight while these are combined
                    tmem_ld_wait();
                    if (ch < 15) {
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) tmem_ld4(lane_addr + gg * 128 + half * 64 + (ch + 1) * 4, v[(ch + 1) & 1][gg]);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const long long u = (long long)v[ch & 1][0][c] * 128 + v[ch & 1][1][c], w = (long long)v[ch & 1][2][c] * 128 + v[ch & 1][3][c];
                        const double du = __longlong_as_double(u + 0x4338000000000000ll) - 6755399441055744.0;
                        const double dw = __longlong_as_double(w + 0x4338000000000000ll) - 6755399441055744.0;
                        x[ch * 4 + c] = fma(du, 16384.0, dw);   // exact: < 2^47
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty);   // TMEM is free: the next pass's MMAs run under phase 2
                if (o.exp & 1) {
                    double sum = 0.0;
#pragma unroll
                    for (int c = 0; c < 64; ++c) sum += x[c];
                    if (sum == 12345.678) g.C[0] = sum;   // keep phase 1 observable
                    continue;
                }
                // phase 2: x * (row scale * column scale) leaves through the warp's own staging (two 32 x 8 images used in turn, so
                // that one reduce can be in flight while the next image is written), 8 columns at a time, as TMA reduce-adds into C
                // (f64 add at the L2; the SM never reads C).  The two passes' additions to an element are applied in a fixed order:
                // pass 0's reduces have been PERFORMED before pass 1 issues its first.
                if (lane == 0) {
                    if (pass == 1) tma_wait_group_all0();
                }
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    const int c0 = half * 64 + cb * 8;
                    double* img = stg + (cb & 1) * (32 * 8);
                    if (lane == 0) tma_wait_group_read1();   // the image written two rounds ago has been read
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double val = x[cb * 8 + c] * (rf * __ldg(o.scB + n0 + c0 + c));
                        img[c * 32 + lane] = (diag && row < c0 + c) ? 0.0 : val;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmC, m0 + quarter * 32, n0 + c0, img);
                        tma_commit_group();
                    }
                }
            }
        }
        if (lane == 0) tma_wait_group_all0();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ==================================================================================================================
// CTA pairs (cta_group::2).  One tcgen05.mma of the pair covers M = 256: two 128-row tiles that share the tile column tj
// (any two tiles of a column segment, they need not be adjacent).  Each CTA stages its OWN row block (8 slices x 128 rows)
// and HALF of the column block (8 slices x 64 rows; the tensor cores read the other half from the peer's shared memory),
// which cuts the shared-memory operand traffic per MMA and SM from 8 KB to 6 KB — the single-CTA kernel is bound by exactly
// that traffic (measured: ~76 B/clk of operand fetch, 105 clk per MMA against 64 ideal).  Roles per CTA as in the single-CTA
// kernel, except: only the leader (cluster rank 0) issues MMAs and commits (multicast to both CTAs' barriers); the peer's
// warp 1 relays "my stage is full" to the leader (remote mbarrier arrive); both CTAs' epilogue warps release the accumulators
// on the leader's barrier.
constexpr int OZ2_STAGES = 4;
constexpr int OZ2_A_BYTES = OZ_PART_BYTES;                       // own row block
constexpr int OZ2_BH_BLOCK = OZ_BLOCK_BYTES / 2;                 // 64 rows x 32 k of one slice
constexpr int OZ2_BH_BYTES = OZ_SLICES * OZ2_BH_BLOCK;           // this CTA's half of the column block
constexpr int OZ2_STAGE_BYTES = OZ2_A_BYTES + OZ2_BH_BYTES;      // 48 KB
constexpr int OZ2_SMEM_BYTES = OZ2_STAGES * OZ2_STAGE_BYTES + OZ_STAGING_BYTES + 1024;
constexpr uint32_t OZ_IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);   // M = 256 over the pair

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void oz_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote arrivals)
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 24)) {
            printf("ozaki_pair_kernel: barrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, a, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_i8_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int PASS>
__device__ __forceinline__ void issue_kstep2(uint32_t tmem_base, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t not_first) {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
        constexpr int G0 = PASS * 4;
#pragma unroll
        for (int i = 0; i <= G0 + gg; ++i) {
            const int j = G0 + gg - i;
            umma_i8_2(tmem_base + gg * 128, desc64(a_lo + i * (OZ_BLOCK_BYTES >> 4), hi), desc64(b_lo + j * (OZ2_BH_BLOCK >> 4), hi), OZ_IDESC2,
                      i > 0 ? 1u : not_first);
        }
    }
}

// pair units: every (band, tile column) segment of the tile enumeration of gemm_tile_decode is cut into pairs of consecutive
// tiles (an odd segment ends with a lone tile: ti1 = -1, the second CTA then shadows the first and stores nothing);
// band_prefix counts UNITS here (oz_pair_plan)
__host__ __device__ inline void oz_pair_decode(const GemmArgs& g, int u, int& ti0, int& ti1, int& tj) {
    const int tm = g.M / GEMM_BM;
    int first, k, seg;
    if (!g.lower) {
        const int ph = (tm + 1) / 2;
        tj = u / ph;
        k = u % ph;
        first = 0;
        seg = tm;
    } else {
        const int PT = g.grp > 0 ? g.grp : 1, S = g.stride > 0 ? g.stride : 1, tn = g.N / GEMM_BN, R = g.band_rows;
        int r = 0;
        while (r + 1 < g.n_bands && g.band_prefix[r + 1] <= u) ++r;
        int o = u - g.band_prefix[r];
        const int lo = g.row_skip + r * R, hi = (lo + R < tm) ? lo + R : tm, h = hi - lo, ph = (h + 1) / 2;
        int nf = (lo / S) * PT + ((lo % S) < PT ? (lo % S) : PT);
        if (nf > tn) nf = tn;
        if (o < nf * ph) {
            const int jl = o / ph;
            tj = (jl / PT) * S + jl % PT;
            first = lo;
            k = o % ph;
            seg = h;
        } else {
            o -= nf * ph;
            int jl = nf;
            for (;; ++jl) {
                tj = (jl / PT) * S + jl % PT;
                const int pc = (hi - tj + 1) / 2;
                if (o < pc) break;
                o -= pc;
            }
            first = tj;
            k = o;
            seg = hi - tj;
        }
    }
    ti0 = first + 2 * k;
    ti1 = (2 * k + 1 < seg) ? ti0 + 1 : -1;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
ozaki_pair_kernel(const __grid_constant__ GemmArgs g, const __grid_constant__ CUtensorMap tmC, const OzArgs o) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    unsigned char* tail = smem + OZ2_STAGES * OZ2_STAGE_BYTES + OZ_STAGING_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);                 // [OZ2_STAGES] this CTA's stage has landed
    uint64_t* empty = full + OZ2_STAGES;                                // [OZ2_STAGES] the pair's MMAs have read the stage (multicast commit)
    uint64_t* peer_full = empty + OZ2_STAGES;                           // [OZ2_STAGES] leader only: the peer's stage has landed (remote arrive)
    uint64_t* tmem_full = peer_full + OZ2_STAGES;                       // accumulators complete (multicast commit)
    uint64_t* tmem_empty = tmem_full + 1;                               // leader only: both CTAs' epilogues have drained them
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1;
    const int u_begin = pair * o.tpc, u_end = min(u_begin + o.tpc, o.tiles);   // o.tiles = pair units of the launch

    if (threadIdx.x == 32) {
        for (int s = 0; s < OZ2_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&peer_full[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * OZ_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers are initialised and their TMEM allocated before anything crosses the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== producer: own row block + own half of the column block =====
        int stage = 0;
        uint32_t phase = 0;
        for (int u = u_begin; u < u_end; ++u) {
            int ti0, ti1, tj;
            oz_pair_decode(g, u, ti0, ti1, tj);
            const int ti = (rank == 1 && ti1 >= 0) ? ti1 : ti0;
            const int8_t* a0 = o.SA + (int64_t)ti * o.KS * OZ_PART_BYTES;
            const int8_t* b0 = o.SB + (int64_t)tj * o.KS * OZ_PART_BYTES + rank * OZ2_BH_BLOCK;
            const int ks0 = g.k_from_tile ? (128 / OZ_KSTEP) * (g.k_tile0 + max(max(ti0, ti1), tj)) : 0;
            for (int pass = 0; pass < 2; ++pass) {
                const int nsl = pass == 0 ? OZ_SLICES / 2 : OZ_SLICES;
                for (int ks = ks0; ks < o.KS; ++ks) {
                    oz_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        unsigned char* dst = ring + stage * OZ2_STAGE_BYTES;
                        if (o.exp & 2) {
                            mbar_arrive(&full[stage]);
                        } else {
                            mbar_arrive_expect_tx(&full[stage], nsl * (OZ_BLOCK_BYTES + OZ2_BH_BLOCK));
                            tma_load_1d(dst, a0 + (int64_t)ks * OZ_PART_BYTES, nsl * OZ_BLOCK_BYTES, &full[stage]);
                            for (int sl = 0; sl < nsl; ++sl)
                                tma_load_1d(dst + OZ2_A_BYTES + sl * OZ2_BH_BLOCK, b0 + (int64_t)ks * OZ_PART_BYTES + sl * OZ_BLOCK_BYTES,
                                            OZ2_BH_BLOCK, &full[stage]);
                        }
                    }
                    __syncwarp();
                    if (++stage == OZ2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0, npass = 0;
        const uint32_t hi = (uint32_t)(oz_desc(0, o.lbo, o.sbo) >> 32), lbo_field = (o.lbo >> 4) << 16;
        const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | lbo_field;
        for (int u = u_begin; u < u_end; ++u) {
            int ks0 = 0;
            if (g.k_from_tile) {
                int ti0, ti1, tj;
                oz_pair_decode(g, u, ti0, ti1, tj);
                ks0 = (128 / OZ_KSTEP) * (g.k_tile0 + max(max(ti0, ti1), tj));
            }
            for (int pass = 0; pass < 2; ++pass, ++npass) {
                if (rank == 0) {
                    oz_wait_cluster(tmem_empty, (npass & 1) ^ 1);   // both epilogues have drained the accumulators
                    tc_fence_after();
                }
                for (int ks = ks0; ks < o.KS; ++ks) {
                    oz_wait(&full[stage], phase);
                    if (rank == 1) {
                        // ===== peer: relay "my stage has landed" to the leader =====
                        if (elect_one()) mbar_arrive_remote(&peer_full[stage], 0);
                    } else {
                        // ===== leader: MMAs of the pair =====
                        oz_wait_cluster(&peer_full[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_lo = ring_lo + stage * (OZ2_STAGE_BYTES >> 4), b_lo = a_lo + (OZ2_A_BYTES >> 4);
                            if (!(o.exp & 4)) {
                                if (pass == 0) issue_kstep2<0>(tmem_base, a_lo, b_lo, hi, ks > ks0);
                                else issue_kstep2<1>(tmem_base, a_lo, b_lo, hi, ks > ks0);
                            }
                            tc_commit2(&empty[stage]);                    // both CTAs' stages are free once these MMAs have read them
                            if (ks == o.KS - 1) tc_commit2(tmem_full);    // both CTAs' accumulators complete
                        }
                    }
                    __syncwarp();
                    if (++stage == OZ2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue (as in the single-CTA kernel, on this CTA's own tile) =====
        const int quarter = warp & 3, row = quarter * 32 + lane, half = (warp - 2) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        double* stg = reinterpret_cast<double*>(smem + OZ2_STAGES * OZ2_STAGE_BYTES) + (warp - 2) * (32 * 16);
        uint32_t npass = 0;
        for (int u = u_begin; u < u_end; ++u) {
            int ti0, ti1, tj;
            oz_pair_decode(g, u, ti0, ti1, tj);
            const bool shadow = rank == 1 && ti1 < 0;   // lone tile: this CTA computes a copy and stores nothing
            const int ti = (rank == 1 && ti1 >= 0) ? ti1 : ti0;
            const bool diag = g.lower && ti == tj;
            const int m0 = ti * 128, n0 = tj * 128;
            const double rs = o.scA[m0 + row];
            for (int pass = 0; pass < 2; ++pass, ++npass) {
                const double rf = rs * o.alpha * (pass == 0 ? 134217728.0 : 0.5);
                oz_wait(tmem_full, npass & 1);
                tc_fence_after();
                double x[64];
                int v[2][4][4];
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) tmem_ld4(lane_addr + gg * 128 + half * 64, v[0][gg]);
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    tmem_ld_wait();
                    if (ch < 15) {
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) tmem_ld4(lane_addr + gg * 128 + half * 64 + (ch + 1) * 4, v[(ch + 1) & 1][gg]);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const long long uu = (long long)v[ch & 1][0][c] * 128 + v[ch & 1][1][c], w = (long long)v[ch & 1][2][c] * 128 + v[ch & 1][3][c];
                        const double du = __longlong_as_double(uu + 0x4338000000000000ll) - 6755399441055744.0;
                        const double dw = __longlong_as_double(w + 0x4338000000000000ll) - 6755399441055744.0;
                        x[ch * 4 + c] = fma(du, 16384.0, dw);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {   // the leader's MMA warp waits for both CTAs' eight epilogue warps
                    if (rank == 0) mbar_arrive(tmem_empty);
                    else mbar_arrive_remote(tmem_empty, 0);
                }
                if (shadow || (o.exp & 1)) {
                    if (o.exp & 1) {
                        double sum = 0.0;
#pragma unroll
                        for (int c = 0; c < 64; ++c) sum += x[c];
                        if (sum == 12345.678) g.C[0] = sum;
                    }
                    continue;
                }
                if (lane == 0 && pass == 1) tma_wait_group_all0();
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const int c0 = half * 64 + cb * 16;
                    if (lane == 0) tma_wait_group_read0();
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const double val = x[cb * 16 + c] * (rf * __ldg(o.scB + n0 + c0 + c));
                        stg[c * 32 + lane] = (diag && row < c0 + c) ? 0.0 : val;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmC, m0 + quarter * 32, n0 + c0, stg);
                        tma_commit_group();
                    }
                }
            }
        }
        if (lane == 0) tma_wait_group_all0();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no MMA of the pair and no remote arrive is in flight any more
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

cudaError_t ozaki_prepare() {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(ozaki_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ozaki_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM_BYTES);
    if (e == cudaSuccess) done = true;
    return e;
}

void ozaki_slice_launch(const double* P, int64_t ld, int64_t rows, int K, int8_t* digits, double* scale, const LaunchCtx& ctx) {
    if (rows <= 0 || K <= 0) return;
    ProfScope ps(ctx, PROF_SLICE, 16.0 * (double)rows * K);   // "flops" slot: bytes moved (8 read + 8 written per element)
    ozaki_rowmax_kernel<<<(unsigned)(rows / 128), 1024, 0, ctx.st>>>(P, ld, K, scale);
    ozaki_digits_kernel<<<dim3((unsigned)(rows / 128), (unsigned)(K / OZ_KSTEP)), 256, 0, ctx.st>>>(P, ld, K / OZ_KSTEP, scale, digits);
}

static int g_oz_exp = 0;
void ozaki_set_experiment(int flags) { g_oz_exp = flags; }

// band rasterisation in PAIR units (oz_pair_decode): the per-band counts of gemm_nt_plan with every (band, column) segment of
// `cnt` tiles contributing ceil(cnt / 2) units; returns the number of units
static int64_t oz_pair_plan(GemmArgs& g) {
    g.band_rows = GEMM_BAND_ROWS;
    g.n_bands = 1;
    g.band_prefix[0] = 0;
    g.band_prefix[1] = 0;
    const int tm = g.M / GEMM_BM, tn = g.N / GEMM_BN;
    if (!g.lower) return (int64_t)tn * ((tm + 1) / 2);
    const int PT = std::max(g.grp, 1), S = std::max(g.stride, 1);
    const int skip = std::min(std::max(g.row_skip, 0), tm);
    int R = GEMM_BAND_ROWS;
    while ((tm + R - 1) / R > GEMM_MAX_BANDS) R *= 2;
    const int nb = std::max(1, (tm - skip + R - 1) / R);
    g.band_rows = R;
    g.n_bands = nb;
    std::vector<int64_t> cnt(nb, 0);
    for (int jl = 0; jl < tn; ++jl) {
        const int tj = (jl / PT) * S + jl % PT;
        if (tj >= tm) break;
        const int first = std::max(tj, skip);
        for (int r = (first - skip) / R; r < nb; ++r) {
            const int lo = std::max(skip + r * R, first), hi = std::min(tm, skip + (r + 1) * R);
            cnt[r] += (hi - lo + 1) / 2;
        }
    }
    int64_t run = 0;
    for (int r = 0; r < nb; ++r) {
        g.band_prefix[r] = (int)run;
        run += cnt[r];
    }
    g.band_prefix[nb] = (int)run;
    return run;
}

int64_t ozaki_update_launch(const GemmArgs& g, const int8_t* digitsA, const double* scaleA, const int8_t* digitsB,
                            const double* scaleB, int tiles_per_cta, const LaunchCtx& ctx, uint32_t lbo, uint32_t sbo) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const int64_t tiles = gemm_nt_tiles(g);
    if (tiles <= 0) return 0;
    // contract: whole tiles / k-steps, int32-exact contraction length, alpha = +-1 (anything else would need a rounding)
    if (g.M % GEMM_BM || g.N % GEMM_BN || g.K % OZ_KSTEP || g.K > 32768 || (g.alpha != 1.0 && g.alpha != -1.0) || !g.beta_one ||
        g.k_upto_col) {
        fprintf(stderr, "libfgp_sm100: ozaki_update_launch: unsupported problem (M=%d N=%d K=%d alpha=%g)\n", g.M, g.N, g.K, g.alpha);
        gemm_nt_flag_error();
        return 0;
    }
    if (ozaki_prepare() != cudaSuccess) {
        gemm_nt_flag_error();
        return 0;
    }
    GemmArgs p = g;
    gemm_nt_plan(p);
    alignas(64) CUtensorMap tmC;
    const int64_t ncols = g.lower ? g.M : g.N;
    // box of the epilogue's reduce-adds: 32 rows x 8 columns (single-CTA kernel) / 32 x 16 (CTA-pair experiment)
    if (!make_tile_map(&tmC, g.C, g.M, ncols, g.ldc, 32, (g_oz_exp & 32) ? 16 : 8)) {
        gemm_nt_flag_error();
        return 0;
    }
    OzArgs o{};
    o.SA = digitsA; o.SB = digitsB; o.scA = scaleA; o.scB = scaleB;
    o.KS = g.K / OZ_KSTEP;
    o.tiles = (int)tiles;
    const int num_sms = gemm_nt_num_sms();
    // tiles per CTA: CTAs are dispatched in waves of one per SM, so the launch lasts about waves x (tpc tiles + set-up); pick the
    // tpc <= 4 (longer-lived CTAs hold back the panel stream: measured, fit 28.8 -> 29.2 ms with up to 8) with the shortest estimate (measured per-tile / per-CTA times; M = 4096: 528 tiles -> 4 per CTA = one wave,
    // 0.138 ms against 0.196 ms with 3), ties to the smaller one so that CTAs retire early for the panel stream
    int tpc = tiles_per_cta;
    if (tpc <= 0) {
        double best = 0.0;
        for (int c = 1; c <= 4; ++c) {
            const int64_t ctas = (tiles + c - 1) / c, waves = (ctas + num_sms - 1) / num_sms;
            const double est = (double)waves * (34.0 * c + 3.0);
            if (c == 1 || est < best * 0.995) {
                best = est;
                tpc = c;
            }
        }
    }
    o.tpc = tpc;
    o.lbo = lbo; o.sbo = sbo;
    o.exp = g_oz_exp;
    o.alpha = g.alpha;
    ProfScope ps(ctx, PROF_TCGEN05, gemm_nt_flops(g));   // f64-equivalent flops; the int8 tensor work is 36 x that
    if (g_oz_exp & 32) {   // CTA pairs (cta_group::2)
        const int64_t units = oz_pair_plan(p);
        o.tiles = (int)units;
        if (tiles_per_cta <= 0) o.tpc = (int)std::max<int64_t>(1, std::min<int64_t>(2, units / num_sms));
        const unsigned grid2 = 2u * (unsigned)((units + o.tpc - 1) / o.tpc);
        ozaki_pair_kernel<<<grid2, OZ_THREADS, OZ2_SMEM_BYTES, ctx.st>>>(p, tmC, o);
        return tiles;
    }
    const unsigned grid = (unsigned)((tiles + tpc - 1) / tpc);
    ozaki_update_kernel<<<grid, OZ_THREADS, OZ_SMEM_BYTES, ctx.st>>>(p, tmC, o);
    return tiles;
}

}  // namespace fgp
