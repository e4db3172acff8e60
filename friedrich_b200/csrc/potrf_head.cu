// potrf_head.cu — the panel HEAD kernel: the critical path of the blocked Cholesky in ONE launch (contract in potrf.cuh).
//
// For a panel of nt <= 4 block columns (128 wide) it factors the (128 nt)^2 diagonal block A11 = L11 L11^T in place and
// forms W = L11^-1 (lower triangular, ld = 512), so that everything BELOW the diagonal block is one tensor-pipe GEMM
// afterwards (L21 = A21 W^T, K = 128 nt) instead of nt dependent diagonal-tile / panel-solve / rank-128-update launch
// triples.  Replaces the per-block-column chain of nalgebra's column loop (Cholesky::new_internal, called at
// src/algebra/mod.rs:83,90) on the part of the matrix every later step waits for.
//
// Grid = 1 + NW CTAs of 256 threads.  Every CTA fits next to ONE resident 64-row GEMM CTA (<= 128 registers per thread,
// ~115 KiB of shared memory), so the launch never has to wait for a whole SM to drain while a trailing update is running.
//   CTA 0       the diagonal tiles, one after the other, entirely in shared memory (head_diag_tile): 4 sub-panels of 32
//               columns — pivot block by one warp with the rows in registers (shuffles only on the dependent chain),
//               rows below by one thread per row, in-tile SYRK on the tensor pipe — and the tile's inverse X = L_jj^-1 IN
//               PLACE over the factor (row block s-1 while the pivot chain of sub-panel s runs).
//   CTAs 1..NW  the tile products between the diagonal tiles, in 32-row slices straight from L2 (DMMA fragments of the B
//               operand by ld.global.cg, the A slice staged in shared memory):
//                   S-phase j   A_ij <- A_ij inv_j^T                       (i > j)
//                   U-phase j   A_ik -= A_ij A_kj^T                        (j < k <= i), tile (j+1, j+1) first
//               and, in the shadow of the next diagonal tile, the off-diagonal blocks of W:
//                   P-phase i   P_ic = sum_{k=c}^{i-1} L_ik W_kc           W-phase i   W_ic = -inv_i P_ic      (c < i)
// The CTAs synchronise through counters in global memory (release / acquire); data written inside the kernel is read back
// with ld.global.cg only.  Waits are acyclic (a phase only waits for earlier phases) and every CTA of the grid becomes
// resident independently of the others, so the kernel cannot deadlock whatever else is running.
#include "potrf.cuh"

namespace fgp {

// Phase timing hook for tools/microbench/head_phases.cu (production builds leave FGP_HEAD_TIMING undefined): clock64 stamps
// of lane 0 of every warp of CTA 0 after each phase function (a stamp after __syncthreads would be taken BEFORE the barrier
// completes: BAR.SYNC.DEFER_BLOCKING only blocks at the next dependent instruction)
#ifdef FGP_HEAD_TIMING
__device__ long long g_head_clk[8 * 64];  // [warp of CTA 0][slot]: each warp's own progress (lane 0)
#define HEAD_MARK(i) do { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) g_head_clk[64 * (threadIdx.x >> 5) + (i)] = clock64(); } while (0)
#else
#define HEAD_MARK(i) do { } while (0)
#endif

namespace {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_LDW = HEAD_PANEL;  // 512: leading dimension of W and of the P scratch
constexpr unsigned FULL = 0xffffffffu;

// ---- shared-memory layout of the diagonal tile: column-major lower TRAPEZOID in bands of 32 columns -----------------------
// band b (columns 32b .. 32b+31) holds rows 32b .. 127 with column stride 132 - 32 b (= height + 4).  Every stride is
// 4 mod 16, which makes both access patterns conflict free: "lane = row, fixed column" (consecutive doubles) and DMMA
// fragments (8 rows x 4 columns: (4 t + g) mod 16 distinct inside a half warp).
__host__ __device__ constexpr int ht_stride(int b) { return 132 - 32 * b; }
__host__ __device__ constexpr int ht_base(int b) { return 4224 * b - 512 * b * (b - 1); }
constexpr int HT_DOUBLES = 10752;
constexpr int HS_LD = 36;                              // scratch blocks of the inverse: S(k, n) at n * 36 + k
constexpr int HEAD_SMEM_DIAG = (HT_DOUBLES + 3 * 32 * HS_LD + 128 + 16 * 20 + 2) * 8;  // tile, S scratch, 1/diagonal, X32 scratch, mbarrier
constexpr int HEAD_AS_LD = 20;                         // staged A slice of a worker task (16 rows): A(r, k) at k * 20 + r
constexpr int HEAD_SMEM_WORK = 384 * HEAD_AS_LD * 8;
constexpr int HEAD_SMEM = HEAD_SMEM_DIAG > HEAD_SMEM_WORK ? HEAD_SMEM_DIAG : HEAD_SMEM_WORK;

__device__ __forceinline__ int ht(int r, int c) {
    const int b = c >> 5;
    return ht_base(b) + (c & 31) * ht_stride(b) + r - 32 * b;
}

// ---- counters -------------------------------------------------------------------------------------------------------
enum { SY_DIAG = 0, SY_S = 1, SY_U = 5, SY_UD = 9, SY_P = 13, SY_W = 17 };

// A wait that outlasts ~2 s of polling can only be a lost dependency (a bug, or a co-operating CTA that died): the kernel
// then reports HEAD_TIMEOUT through `info` and runs on — a loud failure instead of a hung GPU.
__device__ __forceinline__ void head_wait_ge(const int* ctr, int target, int* info) {
    if (threadIdx.x == 0) {
        int v;
        for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if (spins > (1u << 24)) {
                atomicExch(info, HEAD_TIMEOUT);
                break;
            }
            __nanosleep(spins < 64 ? 20 : 100);
        }
    }
    __syncthreads();
}
// every thread's global writes become visible before the counter moves
__device__ __forceinline__ void head_signal(int* ctr, int amount) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && amount > 0) atomicAdd(ctr, amount);
}

__device__ __forceinline__ double head_rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// =====================================================================================================================
// CTA 0: one 128 x 128 diagonal tile
// =====================================================================================================================

// What the scalar fp64 path costs here (tools/microbench/fp64_latency.cu, one warp, dependent issue): DFMA / DMUL 8.4 cycles,
// 64-bit SHFL 26, MUFU.RSQ64H 17, STS->LDS 35, DMMA 26.  Warps issue IN ORDER, so a stream of "operand from another lane
// (SHFL or LDS), then one FMA" runs at the operand latency per FMA unless many operands are fetched ahead — measured: a
// 32-column pivot block with shuffle-broadcast rank-1 updates took 16-33 k cycles, and the thread-per-row substitution
// (496 LDS+FMA pairs) 5-11 k.  Hence: scalar code only for the 8-column pivot chains themselves; every update, every
// triangular solve and every inverse above 8 x 8 is a DMMA product on fragments read straight from the shared-memory tile.

// (a) pivot block of sub-panel s by ONE warp, in four panels of 8 columns.  Lane r owns row r of the block.
//   per column: the dependent chain rs_k -> l_{k+1,k} -> pivot_{k+1} (own registers of lane k+1) -> shuffle -> rsqrt, about 95
//   cycles; beside it the panel's remaining columns get their rank-1 update from the column just stored (broadcast reads);
//   per panel:  the columns behind it inside the block get the panel's rank-8 update on the tensor pipe (<= 6 tiles).
__device__ __noinline__ void head_pivot32(double* T, double* rdiag, int s, int lane, int sub_ok, double sub, int* info,
                                          int col_base) {
    const int S = ht_stride(s);
    double* Tb = T + ht_base(s);  // block element (row i, column k) of the band at Tb[k * S + i]
    const int g = lane >> 2, t = lane & 3;
    int badcol = 1 << 30;
    // zero / negative / NaN pivot (nalgebra: is_zero or try_sqrt fails): substitute if given, else record the column.
    // Pivots below 1e-300 count as zero (the MUFU seed flushes subnormals).
    auto checked = [&](double d, int k) -> double {
        const bool ok = d > 1e-300;
        if (!ok && !sub_ok) badcol = min(badcol, k);
        return ok ? d : (sub_ok ? sub : nan(""));
    };
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const int k0 = 8 * p;
        double a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = Tb[(k0 + c) * S + lane];
        double dk = checked(__shfl_sync(FULL, a[0], k0), k0);
        double rs = head_rsqrt_pos(dk);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = k0 + kk;
            // Only rs and lk are on the dependent chain. The diagonal entry sqrt(dk) = dk * rs + one correction step is needed
            // for the store alone (lane k's own lk feeds nothing: row k is finished).
            const double lk = a[kk] * rs;
            double sk = dk * rs;
            sk = fma(0.5 * rs, fma(-sk, sk, dk), sk);
            if (lane == k) rdiag[32 * s + k] = rs;
            Tb[k * S + lane] = (lane > k) ? lk : ((lane == k) ? sk : 0.0);  // zero above the diagonal: a DMMA operand later
            if (kk < 7) {
                const double pv = fma(-lk, lk, a[kk + 1]);  // the next pivot, exact on lane k+1
                dk = checked(__shfl_sync(FULL, pv, k + 1), k + 1);
                rs = head_rsqrt_pos(dk);
                __syncwarp();
#pragma unroll
                for (int c = kk + 1; c < 8; ++c) a[c] = fma(-lk, Tb[k * S + k0 + c], a[c]);  // L(k0 + c, k): broadcast read
            }
        }
        __syncwarp();
        HEAD_MARK(34 + 2 * p);
        // rank-8 update of the columns behind the panel, inside the block: tiles (mi, ni), p < ni <= mi <= 3.  All six lower
        // tiles are computed in one straight-line block (fragments first, then 12 independent DMMAs); tiles of finished
        // columns (ni <= p) are not stored.  (Nothing is behind the last panel.)
        if (p < 3) {
            double f[3][2], cv[6][2], acc[6][2];
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) f[m][ks] = Tb[(k0 + 4 * ks + t) * S + 8 * (m + 1) + g];  // P(8 (m+1) + g, k0 + 4 ks + t)
            int q = 0;
#pragma unroll
            for (int ni = 1; ni < 4; ++ni)
#pragma unroll
                for (int mi = ni; mi < 4; ++mi, ++q) {
                    cv[q][0] = Tb[(8 * ni + 2 * t) * S + 8 * mi + g];
                    cv[q][1] = Tb[(8 * ni + 2 * t + 1) * S + 8 * mi + g];
                    acc[q][0] = acc[q][1] = 0.0;
                }
            q = 0;
#pragma unroll
            for (int ni = 1; ni < 4; ++ni)
#pragma unroll
                for (int mi = ni; mi < 4; ++mi, ++q)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) dmma884(acc[q][0], acc[q][1], f[mi - 1][ks], f[ni - 1][ks]);
            q = 0;
#pragma unroll
            for (int ni = 1; ni < 4; ++ni)
#pragma unroll
                for (int mi = ni; mi < 4; ++mi, ++q)
                    if (ni > p) {  // (predicated stores: the columns of the panel itself are being read as fragments above)
                        Tb[(8 * ni + 2 * t) * S + 8 * mi + g] = cv[q][0] - acc[q][0];
                        Tb[(8 * ni + 2 * t + 1) * S + 8 * mi + g] = cv[q][1] - acc[q][1];
                    }
        }
        __syncwarp();
        HEAD_MARK(35 + 2 * p);
    }
    if (badcol < 32 && lane == 0) atomicCAS(info, 0, col_base + 32 * s + badcol + 1);
}

// X = L_pp^-1 of the factored 32 x 32 pivot block, IN PLACE, by one warp: the four 8 x 8 diagonal inverses by substitution
// (lane = (block, column): 8 steps), then recursive doubling on the tensor pipe:
//     X16[1,0] = -X8[1] (L[1,0] X8[0])   (two 8x8x8 products per 16-block)      X32[1,0] = -X16[1] (L32[1,0] X16[0])   (two 16x16x16)
// Wx: 16 x 16 scratch, ld 20 (the intermediate product as the next B operand).  The block's strict upper triangle stays zero.
constexpr int HX_LD = 20;
__device__ __noinline__ void head_x32(double* T, double* Wx, const double* rdiag, int s, int lane) {
    const int S = ht_stride(s);
    double* Tb = T + ht_base(s);
    const double* rd = rdiag + 32 * s;
    const int g = lane >> 2, t = lane & 3;
    {   // 8 x 8 diagonal blocks
        const int base = 8 * (lane >> 3), c = lane & 7;
        double x[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) x[r] = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x[r] *= rd[base + r];
#pragma unroll
            for (int r2 = r + 1; r2 < 8; ++r2) x[r2] = fma(-x[r], Tb[(base + r) * S + base + r2], x[r2]);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r >= c) Tb[(base + c) * S + base + r] = x[r];
        __syncwarp();
    }
    HEAD_MARK(32);
    {   // 8 -> 16: both 16-blocks at once (b = 0, 1)
        double sa[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                dmma884(sa[b][0], sa[b][1], Tb[(16 * b + 4 * ks + t) * S + 16 * b + 8 + g],   // L(16b+8+g, 16b+4ks+t)
                        Tb[(16 * b + g) * S + 16 * b + 4 * ks + t]);                             // X8[2b](4ks+t, g)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            Wx[(2 * t) * HX_LD + 8 * b + g] = sa[b][0];
            Wx[(2 * t + 1) * HX_LD + 8 * b + g] = sa[b][1];
        }
        __syncwarp();
        double xa[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                dmma884(xa[b][0], xa[b][1], Tb[(16 * b + 8 + 4 * ks + t) * S + 16 * b + 8 + g],  // X8[2b+1](g, 4ks+t)
                        Wx[g * HX_LD + 8 * b + 4 * ks + t]);                                      // S_b(4ks+t, g)
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            Tb[(16 * b + 2 * t) * S + 16 * b + 8 + g] = -xa[b][0];
            Tb[(16 * b + 2 * t + 1) * S + 16 * b + 8 + g] = -xa[b][1];
        }
        __syncwarp();
    }
    HEAD_MARK(33);
    {   // 16 -> 32
        double sa[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) sa[mi][ni][0] = sa[mi][ni][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double fa[2], fb[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) fa[mi] = Tb[(4 * ks + t) * S + 16 + 8 * mi + g];   // L(16+8mi+g, 4ks+t)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) fb[ni] = Tb[(8 * ni + g) * S + 4 * ks + t];        // X16[0](4ks+t, 8ni+g)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884(sa[mi][ni][0], sa[mi][ni][1], fa[mi], fb[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) Wx[(8 * ni + 2 * t + e) * HX_LD + 8 * mi + g] = sa[mi][ni][e];
        __syncwarp();
        double xa[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) xa[mi][ni][0] = xa[mi][ni][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double fa[2], fb[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) fa[mi] = Tb[(16 + 4 * ks + t) * S + 16 + 8 * mi + g];  // X16[1](8mi+g, 4ks+t)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) fb[ni] = Wx[(8 * ni + g) * HX_LD + 4 * ks + t];        // S(4ks+t, 8ni+g)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884(xa[mi][ni][0], xa[mi][ni][1], fa[mi], fb[ni]);
        }
        __syncwarp();
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) Tb[(8 * ni + 2 * t + e) * S + 16 + 8 * mi + g] = -xa[mi][ni][e];
        __syncwarp();
    }
}

// (b) rows below the pivot block: A_below <- A_below X^T with X = L_pp^-1 (already in place of L_pp), on the tensor pipe.
// Unit = 8 rows x 32 columns (all of the unit's A fragments are read before anything is written: in place is safe);
// X(n, k) = 0 for k > n, so tile column ni contracts over k < 8 (ni + 1).  Units [u_begin, u_end) are shared by the nw warps
// of the calling group (gw = index inside it); unit u covers band rows 32 + 8 u ..: units 0..3 are the NEXT pivot block's rows.
__device__ __noinline__ void head_rows_below_mma(double* T, int s, int u_begin, int u_end, int gw, int nw, int lane) {
    const int S = ht_stride(s);
    double* Tb = T + ht_base(s);
    const int g = lane >> 2, t = lane & 3;
    for (int u = u_begin + gw; u < u_end; u += nw) {
        const int m0 = 32 + 8 * u;
        double fa[8], acc[4][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) fa[ks] = Tb[(4 * ks + t) * S + m0 + g];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                if (ks < 2 * ni + 2) dmma884(acc[ni][0], acc[ni][1], fa[ks], Tb[(4 * ks + t) * S + 8 * ni + g]);  // X(8ni+g, 4ks+t)
        __syncwarp();
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            Tb[(8 * ni + 2 * t) * S + m0 + g] = acc[ni][0];
            Tb[(8 * ni + 2 * t + 1) * S + m0 + g] = acc[ni][1];
        }
    }
}

// (c) in-tile SYRK on the tensor pipe: A22 -= X X^T, X = rows below of sub-panel s.  Unit q = 16 rows x 16 columns of the
// 32-block blk = q / 4 (blocks (bi, bj <= bi) of the trailing part, enumerated row by row: block 0 is the NEXT pivot block),
// quarter (hr, hf) = (row half, column half); the strictly upper quarter of a diagonal block is skipped, and so are its 8 x 8
// tiles above the diagonal.  Units [q_begin, q_end) are shared by the nw warps of the calling group.
__device__ __noinline__ void head_syrk(double* T, int s, int q_begin, int q_end, int gw, int nw, int lane) {
    const int S = ht_stride(s);
    const double* Tb = T + ht_base(s);
    const int g = lane >> 2, t = lane & 3;
    for (int q = q_begin + gw; q < q_end; q += nw) {
        const int blk = q >> 2, hr = (q >> 1) & 1, hf = q & 1;
        int bi = 0, rem = blk;
        while (rem > bi) { rem -= bi + 1; ++bi; }  // blk = bi (bi + 1) / 2 + bj
        const int bj = rem;
        const bool dg = (bi == bj);
        if (dg && hr == 0 && hf == 1) continue;
        const int m0 = 32 + 32 * bi + 16 * hr, n0 = 32 + 32 * bj + 16 * hf;  // band-relative rows of the two operands
        double acc[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const bool skip01 = dg && hr == hf;  // diagonal quarter: tile (mi = 0, ni = 1) lies above the diagonal
#pragma unroll
        for (int kk = 0; kk < 32; kk += 4) {
            double fa[2], fb[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) fa[mi] = Tb[(kk + t) * S + m0 + 8 * mi + g];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) fb[ni] = Tb[(kk + t) * S + n0 + 8 * ni + g];
            dmma884(acc[0][0][0], acc[0][0][1], fa[0], fb[0]);
            if (!skip01) dmma884(acc[0][1][0], acc[0][1][1], fa[0], fb[1]);
            dmma884(acc[1][0][0], acc[1][0][1], fa[1], fb[0]);
            dmma884(acc[1][1][0], acc[1][1][1], fa[1], fb[1]);
        }
        // C(row, col): tile row 32 s + m0 + .., tile column 32 s + n0 + .. = band s + 1 + bj
        const int bc = s + 1 + bj, Sc = ht_stride(bc);
        double* Tc = T + ht_base(bc) + 32 * (bi - bj) + 16 * hr;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) Tc[(16 * hf + 8 * ni + 2 * t + e) * Sc + 8 * mi + g] -= acc[mi][ni][e];
    }
}

__device__ __forceinline__ void head_group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Off-diagonal blocks of row block i (i >= 1) of the tile's inverse X = L^-1, in place over row block i of L, in two halves run
// by a group of nw warps (gw = index inside the group); X_ii is already in place (head_x32):
//     sums:   L_ik (k < i) -> global memory (final, about to be overwritten);  S_ij = sum_{k=j}^{i-1} L_ik X_kj -> Ws[j]   (j < i)
//     apply:  X_ij = -X_ii S_ij   over L_ij                                  (group barrier between the two)
__device__ __noinline__ void head_inverse_sums(double* T, double* Ws, int i, int gw, int nw, int lane, double* __restrict__ Ag,
                                               int64_t lda) {
    const int g = lane >> 2, t = lane & 3;
    for (int c = gw; c < 32 * i; c += nw) Ag[32 * i + lane + (int64_t)c * lda] = T[ht(32 * i + lane, c)];
    for (int u = gw; u < 2 * i; u += nw) {  // unit (j, half): 32 x 16 block of S_ij, K = 32 (i - j)
        const int j = u >> 1, n0 = 16 * (u & 1);
        double acc[4][2][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const int Sj = ht_stride(j);
        for (int k = j; k < i; ++k) {
            const int Sk = ht_stride(k);
            const double* La = T + ht_base(k) + 32 * (i - k);           // L_ik(m, kk) at La[kk * Sk + m]
            const double* Xb = T + ht_base(j) + 32 * (k - j);           // X_kj(kk, n) at Xb[n * Sj + kk]
#pragma unroll
            for (int kk = 0; kk < 32; kk += 4) {
                double fa[4], fb[2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) fa[mi] = La[(kk + t) * Sk + 8 * mi + g];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) fb[ni] = Xb[(n0 + 8 * ni + g) * Sj + kk + t];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
            }
        }
        double* Sd = Ws + j * 32 * HS_LD;  // S_ij(m, n) at Sd[n * 36 + m]
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) Sd[(n0 + 8 * ni + 2 * t + e) * HS_LD + 8 * mi + g] = acc[mi][ni][e];
    }
}

__device__ __noinline__ void head_inverse_apply(double* T, const double* Ws, int i, int gw, int nw, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int Si = ht_stride(i);
    const double* Xa = T + ht_base(i);  // X_ii(m, kk) at Xa[kk * Si + m], zeros above the diagonal
    for (int u = gw; u < 2 * i; u += nw) {  // unit (j, half), K = 32
        const int j = u >> 1, n0 = 16 * (u & 1);
        double acc[4][2][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const double* Sd = Ws + j * 32 * HS_LD;
#pragma unroll
        for (int kk = 0; kk < 32; kk += 4) {
            double fa[4], fb[2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) fa[mi] = Xa[(kk + t) * Si + 8 * mi + g];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) fb[ni] = Sd[(n0 + 8 * ni + g) * HS_LD + kk + t];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
        }
        const int Sj = ht_stride(j);
        double* Xd = T + ht_base(j) + 32 * (i - j);  // X_ij(m, n) at Xd[n * Sj + m]
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) Xd[(n0 + 8 * ni + 2 * t + e) * Sj + 8 * mi + g] = -acc[mi][ni][e];
    }
}

// 1-D bulk copies between the tile in shared memory and global memory (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void head_bulk_store(double* gdst, const double* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// One diagonal tile: Ag (global, ld = lda) -> L in place, X = L^-1 -> inv (128 x 128, ld 128); the strict upper triangle of the
// destination is zero from allocation and only zeros are ever written there.  All 256
// threads; `bar` = mbarrier for the tile load, `parity` = its phase (tile index & 1).
//
// Warp roles while warp 0 runs the pivot chain of sub-panel s (and the pivot block's inverse): warps 1,2,3,5,6,7 compute the
// off-diagonal inverse blocks of row block s-1.  Warp 4 stays idle: it shares warp 0's scheduler and fp64 pipe, and DMMA work
// there doubles the pivot chain's time (measured: 330 vs 168 cycles per column).
__device__ void head_diag_tile(double* smem, uint64_t* bar, uint32_t parity, double* __restrict__ Ag, int64_t lda,
                               double* __restrict__ inv, int has_sub, double sub, int* info, int col_base) {
    double* T = smem;
    double* Ws = smem + HT_DOUBLES;
    double* rdiag = Ws + 3 * 32 * HS_LD;
    double* Wx = rdiag + 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    HEAD_MARK(0);
    // Load: column c from its even row on or above the diagonal (16-byte alignment) to row 127 in ONE bulk copy; the entries
    // above the diagonal inside the 32 x 32 diagonal blocks are whatever the source / the previous tile left — every one of
    // them is overwritten with zero by the pivot step of its column before anything reads it as an operand.
    asm volatile("fence.proxy.async;\n" ::: "memory");  // generic accesses (previous tile's T, the workers' global writes) before the TMA's
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(bar, 8320 * 8);  // sum over columns of 128 - (c & ~1)
    if (tid < 128) {
        const int c = tid, r0 = c & ~1;
        tma_load_1d(T + ht(r0, c), Ag + r0 + (int64_t)c * lda, (uint32_t)(128 - r0) * 8, bar);
    }
    mbar_wait(bar, parity);
    HEAD_MARK(1);
    const int sub_ok = has_sub && sub > 0.0;
    const int gw = (warp < 4) ? warp - 1 : warp - 2;  // index of warps 1,2,3,5,6,7 in the trailing group (warps 0 and 4: not in it)
    const bool grp = (warp != 0 && warp != 4);
    // pivot block s by warp 0: factor, store to global memory (zeros above the diagonal included) before the inverse replaces it
    auto pivot_and_invert = [&](int s) {
        head_pivot32(T, rdiag, s, lane, sub_ok, sub, info, col_base);
        HEAD_MARK(2 + 6 * s);
        const double* Tb = T + ht_base(s);
        const int S = ht_stride(s);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
            double v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = Tb[(c0 + c) * S + lane];
#pragma unroll
            for (int c = 0; c < 8; ++c) Ag[32 * s + lane + (int64_t)(32 * s + c0 + c) * lda] = v[c];
        }
        __syncwarp();
        head_x32(T, Wx, rdiag, s, lane);
        HEAD_MARK(3 + 6 * s);
    };
    if (warp == 0) pivot_and_invert(0);
    __syncthreads();
    // In-tile look-ahead: after sub-panel s only what the NEXT pivot block needs is done by everybody (A: its 32 rows of the
    // rows-below solve, B: its own SYRK block); then warp 0 runs the next pivot chain while the trailing group finishes
    // sub-panel s (the other rows below, the other SYRK blocks) and advances the tile's inverse.
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
        const int nu = (96 - 32 * s) / 8, nbk = 3 - s, nq = 2 * nbk * (nbk + 1);
        HEAD_MARK(4 + 6 * s);
        if (warp < 4) head_rows_below_mma(T, s, 0, 4, warp, 4, lane);             // A
        __syncthreads();
        HEAD_MARK(5 + 6 * s);
        if (warp < 4) head_syrk(T, s, 0, 4, warp, 4, lane);                       // B
        __syncthreads();
        HEAD_MARK(6 + 6 * s);
        if (warp == 0) {
            pivot_and_invert(s + 1);                                              // C, warp 0
        } else if (grp) {                                                         // C, trailing group
            head_rows_below_mma(T, s, 4, nu, gw, 6, lane);
            head_group_bar(1, 192);
            head_syrk(T, s, 4, nq, gw, 6, lane);
            head_inverse_apply(T, Ws, s, gw, 6, lane);   // S_sj was formed one round earlier (no-op for s = 0)
            head_group_bar(1, 192);
            head_inverse_sums(T, Ws, s + 1, gw, 6, lane, Ag, lda);  // L_{s+1,k} and X_kj, k <= s, are final
            HEAD_MARK(7 + 6 * s);
        }
        __syncthreads();
    }
    HEAD_MARK(26);
    head_inverse_apply(T, Ws, 3, warp, 8, lane);  // S_3j is ready: only the last row block's products are exposed
    HEAD_MARK(27);
    // X (lower trapezoid columns, zeros above the diagonal inside the diagonal blocks) -> inv and W_jj by bulk stores
    fence_proxy_async_smem();
    __syncthreads();
    if (tid < 128) {
        const int c = tid, b = c >> 5;
        const uint32_t bytes = (uint32_t)(128 - 32 * b) * 8;
        head_bulk_store(inv + 32 * b + c * 128, T + ht(32 * b, c), bytes);  // (the workers copy inv_j into W_jj)
        tma_commit_group();
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // the writes have been performed (not only the reads)
    }
    HEAD_MARK(28);
}

// =====================================================================================================================
// worker CTAs: C(16 x 128 slice) = alpha * (C0 / alpha + A(16 x K) * op(B)), B by fragments straight from L2
// =====================================================================================================================
// One SM delivers 128 fp64 flop per clock whatever the instruction mix, so a 128^3 tile product is 16.8 us of ONE SM: the
// products between two diagonal tiles are cut into 16-row slices (0.5 MFlop = 2.1 us each) and spread over up to 48 CTAs.
// BNN = false: B(n, k) at Bg[n + k ldb] (NT product)      BNN = true: B(k, n) at Bg[k + n ldb] (NN product)
// kw: this warp's contraction length (multiple of 32, <= K): operands known to be zero beyond it are skipped.
constexpr int HEAD_SR = 16;                 // rows of a worker slice
constexpr int HEAD_SLICES = TILE / HEAD_SR; // slices per tile
template <bool BNN>
__device__ __forceinline__ void head_task(double* As, const double* Ag, int64_t lda, int K, const double* Bg, int64_t ldb,
                                          double* Cg, int64_t ldc, double alpha, bool accumulate, int kw) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    // stage the A slice: As[k * 20 + r]
    for (int base = tid; base < HEAD_SR * K; base += 8 * HEAD_THREADS) {  // 8 independent L2 loads in flight per thread
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * HEAD_THREADS;
            v[u] = (idx < HEAD_SR * K) ? __ldcg(Ag + (idx & (HEAD_SR - 1)) + (int64_t)(idx / HEAD_SR) * lda) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * HEAD_THREADS;
            if (idx < HEAD_SR * K) As[(idx / HEAD_SR) * HEAD_AS_LD + (idx & (HEAD_SR - 1))] = v[u];
        }
    }
    const int n0 = 16 * warp;
    double acc[2][2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                acc[mi][ni][e] = accumulate ? alpha * __ldcg(Cg + (8 * mi + g) + (int64_t)(n0 + 8 * ni + 2 * t + e) * ldc) : 0.0;
    const double* Bw = BNN ? (Bg + t + (int64_t)(n0 + g) * ldb) : (Bg + (n0 + g) + (int64_t)t * ldb);
    const int64_t bk = BNN ? 1 : ldb;       // step of one k
    const int64_t bn = BNN ? 8 * ldb : 8;   // step of one 8-column fragment
    // B fragments: a ring of 8 k-steps (32 k); the fragment of k-step ks + 8 is requested right after k-step ks has used its
    // slot, i.e. one ring revolution (32 DMMA per warp, >= 512 pipe cycles) ahead of its use
    double bf[8][2];
    const int nks = kw / 4;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) bf[u][ni] = (u < nks) ? __ldcg(Bw + (int64_t)(4 * u) * bk + ni * bn) : 0.0;
    __syncthreads();
    for (int ks0 = 0; ks0 < nks; ks0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ks = ks0 + u;
            const double* Ak = As + (4 * ks + t) * HEAD_AS_LD + g;
            const double fa0 = Ak[0], fa1 = Ak[8];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                dmma884(acc[0][ni][0], acc[0][ni][1], fa0, bf[u][ni]);
                dmma884(acc[1][ni][0], acc[1][ni][1], fa1, bf[u][ni]);
            }
            if (ks + 8 < nks) {
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) bf[u][ni] = __ldcg(Bw + (int64_t)(4 * (ks + 8)) * bk + ni * bn);
            }
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) Cg[(8 * mi + g) + (int64_t)(n0 + 8 * ni + 2 * t + e) * ldc] = alpha * acc[mi][ni][e];
    __syncthreads();  // the staged slice is free again
}

// dst(16 x 128 slice) = src : the diagonal block of W is the tile's inverse
__device__ __forceinline__ void head_copy_task(const double* src, int64_t lds, double* dst, int64_t ldd) {
    const int r = threadIdx.x & (HEAD_SR - 1), c0 = threadIdx.x / HEAD_SR;
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + r + (int64_t)(c0 + 16 * u) * lds);
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[r + (int64_t)(c0 + 16 * u) * ldd] = v[u];
}

struct HeadArgs {
    double* A;        // element (0, 0) of the panel's diagonal block, ld = lda
    int64_t lda;
    int nt;           // 128-tiles in the block (1..4)
    double* inv;      // [nt][128*128]
    double* W;        // 512 x 512 (ld 512), lower block triangle written
    double* P;        // 512 x 512 scratch
    int* sync;        // 32 ints, zero on entry
    int has_sub;
    double sub;
    int* info;
    int col_base;
};

__global__ void __launch_bounds__(HEAD_THREADS, 2) potrf_head_kernel(const HeadArgs a) {
    extern __shared__ __align__(16) double head_smem[];
    const int nt = a.nt;
    const int64_t lda = a.lda;
    if (blockIdx.x == 0) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(head_smem + HT_DOUBLES + 3 * 32 * HS_LD + 128 + 16 * 20);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        for (int j = 0; j < nt; ++j) {
            if (j > 0) head_wait_ge(a.sync + SY_UD + (j - 1), HEAD_SLICES, a.info);
            double* Ajj = a.A + (int64_t)j * TILE * (lda + 1);
            head_diag_tile(head_smem, bar, (uint32_t)(j & 1), Ajj, lda, a.inv + (int64_t)j * TILE * TILE, a.has_sub, a.sub, a.info,
                           a.col_base + j * TILE);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sync + SY_DIAG), "r"(j + 1) : "memory");
        }
        return;
    }
    const int w = blockIdx.x - 1, NW = gridDim.x - 1, warp = threadIdx.x >> 5;
    constexpr int SL = HEAD_SLICES;
    double* As = head_smem;
    for (int j = 0; j < nt; ++j) {
        head_wait_ge(a.sync + SY_DIAG, j + 1, a.info);
        const double* invj = a.inv + (int64_t)j * TILE * TILE;
        // ---- S-phase j: A_ij <- A_ij inv_j^T, i = j+1 .. nt-1, SL slices each; inv_j(n, k) = 0 for k > n
        const int nS = SL * (nt - 1 - j);
        int mine = 0;
        for (int tsk = w; tsk < nS; tsk += NW, ++mine) {
            const int i = j + 1 + tsk / SL, sl = tsk % SL;
            double* Aij = a.A + (int64_t)(i * TILE + HEAD_SR * sl) + (int64_t)j * TILE * lda;
            head_task<false>(As, Aij, lda, TILE, invj, TILE, Aij, lda, 1.0, false, 32 * (warp / 2 + 1));
        }
        head_signal(a.sync + SY_S + j, mine);
        if (nS > 0) {
            head_wait_ge(a.sync + SY_S + j, nS, a.info);
            // ---- U-phase j: A_ik -= A_ij A_kj^T for j < k <= i < nt; tile (j+1, j+1) first (the next diagonal tile waits for it)
            const int m = nt - 1 - j, nU = SL * (m * (m + 1) / 2);
            mine = 0;
            for (int tsk = w; tsk < nU; tsk += NW, ++mine) {
                const int tile = tsk / SL, sl = tsk % SL;
                int k = 0, rem = tile;  // tiles enumerated column by column: (j+1,j+1), (j+2,j+1), .., (nt-1,j+1), (j+2,j+2), ..
                while (rem >= m - k) { rem -= m - k; ++k; }
                const int kk = j + 1 + k, i = kk + rem;
                const double* Aij = a.A + (int64_t)(i * TILE + HEAD_SR * sl) + (int64_t)j * TILE * lda;
                const double* Akj = a.A + (int64_t)kk * TILE + (int64_t)j * TILE * lda;
                double* Cik = a.A + (int64_t)(i * TILE + HEAD_SR * sl) + (int64_t)kk * TILE * lda;
                head_task<false>(As, Aij, lda, TILE, Akj, lda, Cik, lda, -1.0, true, TILE);
                if (tile == 0) head_signal(a.sync + SY_UD + j, 1);
            }
            head_signal(a.sync + SY_U + j, mine);
        }
        // ---- in the shadow of the next diagonal tile: block row j of W (W_jj = inv_j, W_jc = -inv_j P_jc), then the P sums
        //      of block row j+1
        {
            if (j >= 1) head_wait_ge(a.sync + SY_P + j, SL * j, a.info);
            const int nW = SL * (j + 1);
            mine = 0;
            for (int tsk = w; tsk < nW; tsk += NW, ++mine) {
                const int c = tsk / SL, sl = tsk % SL;
                double* Wjc = a.W + (int64_t)(j * TILE + HEAD_SR * sl) + (int64_t)c * TILE * HEAD_LDW;
                if (c == j) {
                    head_copy_task(invj + HEAD_SR * sl, TILE, Wjc, HEAD_LDW);
                } else {  // inv_j(m, k) = 0 for k > m: the slice's rows end at 16 sl + 15
                    const double* Pjc = a.P + (int64_t)j * TILE + (int64_t)c * TILE * HEAD_LDW;
                    const int kw = 32 * ((HEAD_SR * (sl + 1) + 31) / 32);
                    head_task<true>(As, invj + HEAD_SR * sl, TILE, kw, Pjc, HEAD_LDW, Wjc, HEAD_LDW, -1.0, false, kw);
                }
            }
            head_signal(a.sync + SY_W + j, mine);
        }
        if (j + 1 < nt) {
            head_wait_ge(a.sync + SY_W + j, SL * (j + 1), a.info);
            const int i = j + 1;
            mine = 0;
            for (int tsk = w; tsk < SL * i; tsk += NW, ++mine) {  // P_ic = sum_{k=c}^{i-1} L_ik W_kc, K = 128 (i - c)
                const int c = tsk / SL, sl = tsk % SL;
                const double* Lic = a.A + (int64_t)(i * TILE + HEAD_SR * sl) + (int64_t)c * TILE * lda;
                const double* Wcc = a.W + (int64_t)c * TILE * (HEAD_LDW + 1);
                double* Pic = a.P + (int64_t)(i * TILE + HEAD_SR * sl) + (int64_t)c * TILE * HEAD_LDW;
                head_task<true>(As, Lic, lda, TILE * (i - c), Wcc, HEAD_LDW, Pic, HEAD_LDW, 1.0, false, TILE * (i - c));
            }
            head_signal(a.sync + SY_P + i, mine);
            // every U task of this step must have landed before the next S-phase reads the updated tiles
            const int m = nt - 1 - j;
            head_wait_ge(a.sync + SY_U + j, SL * (m * (m + 1) / 2), a.info);
        }
    }
}

// inv tile -> its transpose (the adjoint solves read invT); all tiles of a factorisation in one launch, off the critical path
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const double* __restrict__ inv, double* __restrict__ invT) {
    __shared__ double tile[32][33];
    const double* src = inv + (int64_t)blockIdx.z * TILE * TILE;
    double* dst = invT + (int64_t)blockIdx.z * TILE * TILE;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(bx + tx) + (by + r) * TILE];   // element (row bx+tx, col by+r)
    __syncthreads();
    for (int r = ty; r < 32; r += 8) dst[(by + tx) + (bx + r) * TILE] = tile[tx][r];   // dst(row by+tx, col bx+r) = src(bx+r, by+tx)
}

}  // namespace

cudaError_t potrf_head_prepare() {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(potrf_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM);
    if (e == cudaSuccess) done = true;
    return e;
}

// one round of the widest phase: U-phase 0 has 8 slices x (nt-1) nt / 2 tiles; a single tile still gets W_00 copied by 8 CTAs
int potrf_head_workers(int nt) { return nt <= 1 ? 8 : (nt == 2 ? 16 : (nt == 3 ? 24 : 48)); }

void launch_potrf_head(double* A, int64_t lda, int nt, double* inv, double* W, double* P, int* sync, int has_sub, double sub,
                       int* info, int col_base, const LaunchCtx& c) {
    HeadArgs a{A, lda, nt, inv, W, P, sync, has_sub, sub, info, col_base};
    const double t = (double)nt * TILE;
    ProfScope ps(c, PROF_POTRF_DIAG, 2.0 * t * t * t / 3.0);  // Cholesky of the block + its triangular inverse
    potrf_head_kernel<<<1 + potrf_head_workers(nt), HEAD_THREADS, HEAD_SMEM, c.st>>>(a);
}

void launch_transpose_tiles(const double* inv, double* invT, int64_t nb, cudaStream_t st) {
    if (nb <= 0) return;
    transpose_tiles_kernel<<<dim3(4, 4, (unsigned)nb), 256, 0, st>>>(inv, invT);
}

}  // namespace fgp
