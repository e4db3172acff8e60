// potrf.cu — diagonal-tile factorisation kernel and the blocked lower Cholesky driver (contract in potrf.cuh).
#include "potrf.cuh"
#include "ozaki.cuh"

namespace fgp {

// Phase timing hook for tools/microbench/diag_phases.cu (production builds leave FGP_DIAG_TIMING undefined)
#ifdef FGP_DIAG_TIMING
__device__ long long g_diag_clk[32];
#define DIAG_MARK(i) do { if (threadIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define DIAG_MARK(i) do { } while (0)
#endif

// 1/sqrt(x) for a positive, normal x without the library routine's special-case branches (they would split the pivot
// chain into basic blocks): MUFU.RSQ64H seed (~2^-22) + one third-order step y (1 + e/2 + 3e^2/8), e = 1 - x y^2 -> <= 1 ulp.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// In-tile SYRK of the diagonal-tile factorisation: A[r][c] -= sum_k X[r][k] X[c][k] over the TN x TN block that follows
// the 32-column sub-panel at column c0 (X = the sub-panel's rows below its pivot block). Register tiled: warp w owns rows
// {w + 16 a}, lane l owns columns {l + 32 b}; row operands are warp-wide broadcasts, column operands are conflict free
// (row stride 129). Pairs (a, b) entirely above the diagonal are skipped (warp-uniform test).
template <int TN>
__device__ __forceinline__ void diag_syrk(double* T, int c0, int warp, int lane) {
    constexpr int NA = TN / 16, NB = TN / 32;
    const int base = c0 + 32;
    double acc[NA][NB];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b] = 0.0;
    const double* xr = T + (base + warp) * DIAG_DS + c0;
    const double* xc = T + (base + lane) * DIAG_DS + c0;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        double fr[NA], fc[NB];
#pragma unroll
        for (int a = 0; a < NA; ++a) fr[a] = xr[16 * a * DIAG_DS + k];
#pragma unroll
        for (int b = 0; b < NB; ++b) fc[b] = xc[32 * b * DIAG_DS + k];
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (warp + 16 * a + 1 > 32 * b) acc[a][b] = fma(fr[a], fc[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int r = warp + 16 * a, c = lane + 32 * b;
            if (r >= c) T[(base + r) * DIAG_DS + base + c] -= acc[a][b];
        }
}

// ---- X = L^-1 of the diagonal tile, in 32-blocks.  X[r][c] (r > c) lives TRANSPOSED at T[c*DS + r] (the strict upper
// triangle of T is otherwise unused), X[r][r] = rdiag[r].  Row block i:  X_ii = L_ii^-1,  X_ij = -X_ii S_j with
// S_j = sum_{k=j}^{i-1} L_ik X_kj  (j < i).  It needs block row i of L and the row blocks < i of X, so row block s-1 is
// computed by the otherwise idle warps WHILE warp 0 runs the pivot chain of sub-panel s; only row block 3 is exposed.
// A 32x32 output block is split over 4 warps (8 rows each); lane (rg, cg) = (lane >> 3, lane & 7) owns rows
// {rg, rg + 4} of the slice and columns {cg + 8 b}: per contraction index 2 + 4 shared-memory loads feed 8 FMAs.

// X_ii by one warp: lane c owns column c of the block's inverse (forward substitution on e_c)
__device__ __forceinline__ void diag_inv_block(double* T, const double* rdiag, int i, int lane) {
    const int c0 = 32 * i;
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        x[r] *= rdiag[c0 + r];
#pragma unroll
        for (int r2 = r + 1; r2 < 32; ++r2) x[r2] = fma(-x[r], T[(c0 + r2) * DIAG_DS + c0 + r], x[r2]);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r)
        if (r > lane) T[(c0 + lane) * DIAG_DS + c0 + r] = x[r];
}

// rows [8q, 8q+8) of S_j -> W[j][r][c]
__device__ __forceinline__ void diag_inv_sum(const double* T, double* W, const double* rdiag, int i, int j, int q, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    double acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const double* lrow = T + (32 * i + 8 * q + rg) * DIAG_DS;  // L[32 i + 8 q + rg (+4)][.]
    const double* xcol = T + (32 * j + cg) * DIAG_DS;          // X[.][32 j + cg (+8 b)], transposed storage
    // k = j: X_jj is lower triangular with its diagonal in rdiag
#pragma unroll 4
    for (int m = 0; m < 32; ++m) {
        const double p0 = lrow[32 * j + m], p1 = lrow[4 * DIAG_DS + 32 * j + m];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = cg + 8 * b;
            double xv = xcol[8 * b * DIAG_DS + 32 * j + m];
            xv = (m > c) ? xv : ((m == c) ? rdiag[32 * j + c] : 0.0);
            acc[0][b] = fma(p0, xv, acc[0][b]);
            acc[1][b] = fma(p1, xv, acc[1][b]);
        }
    }
    for (int k = j + 1; k < i; ++k) {
#pragma unroll 4
        for (int m = 0; m < 32; ++m) {
            const double p0 = lrow[32 * k + m], p1 = lrow[4 * DIAG_DS + 32 * k + m];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double xv = xcol[8 * b * DIAG_DS + 32 * k + m];
                acc[0][b] = fma(p0, xv, acc[0][b]);
                acc[1][b] = fma(p1, xv, acc[1][b]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) W[(j * 32 + 8 * q + rg + 4 * a) * 33 + cg + 8 * b] = acc[a][b];
}

// rows [8q, 8q+8) of X_ij = -X_ii S_j -> transposed storage
__device__ __forceinline__ void diag_inv_apply(double* T, const double* W, const double* rdiag, int i, int j, int q, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    const int r0 = 8 * q + rg, r1 = r0 + 4;
    double acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const double* xi = T + 32 * i * DIAG_DS + 32 * i;  // X_ii[r][m] (m < r) at xi[m * DS + r]
    const double* sb = W + j * 32 * 33 + cg;
    const double d0 = rdiag[32 * i + r0], d1 = rdiag[32 * i + r1];
#pragma unroll 4
    for (int m = 0; m < 8 * q + 8; ++m) {  // X_ii[r][m] = 0 for m > r: rows of this slice end at 8q + 7
        double p0 = xi[m * DIAG_DS + r0], p1 = xi[m * DIAG_DS + r1];
        p0 = (m < r0) ? p0 : ((m == r0) ? d0 : 0.0);
        p1 = (m < r1) ? p1 : ((m == r1) ? d1 : 0.0);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double sv = sb[m * 33 + 8 * b];
            acc[0][b] = fma(p0, sv, acc[0][b]);
            acc[1][b] = fma(p1, sv, acc[1][b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        T[(32 * j + cg + 8 * b) * DIAG_DS + 32 * i + r0] = -acc[0][b];
        T[(32 * j + cg + 8 * b) * DIAG_DS + 32 * i + r1] = -acc[1][b];
    }
}

// Row block i of X by the warp group gw = 0 .. ng-1 (ng >= 4 i + 1), synchronised on named barrier `bar_id`.
__device__ __forceinline__ void diag_inv_rowblock(double* T, double* W, const double* rdiag, int i, int gw, int ng, int lane,
                                                  int bar_id) {
    if (gw == 0) diag_inv_block(T, rdiag, i, lane);
    else if (gw <= 4 * i) diag_inv_sum(T, W, rdiag, i, (gw - 1) >> 2, (gw - 1) & 3, lane);
    if (i == 0) return;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * ng) : "memory");
    if (gw < 4 * i) diag_inv_apply(T, W, rdiag, i, gw >> 2, gw & 3, lane);
}

// One CTA, 512 threads: factor the 128x128 diagonal tile in shared memory, write L back, invert it and write the inverse
// and its transpose (`inv`, `invT`, ld = 128, other triangle zeroed).
//   factorisation: 4 sub-panels of 32 columns — (a) the 32x32 pivot block by ONE warp (lane r owns row r; branch-free
//   pivots with a fast rsqrt; rank-1 updates are applied at once only inside the current group of 8 columns, the columns to
//   the right get their 8 updates in one batch, which keeps the 128-long dependent pivot chain short), (b) the rows below
//   by one thread per row, (c) the in-tile SYRK by all warps.  The inverse is computed by the idle warps during (a).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ inv, double* __restrict__ invT, int has_sub,
                  double sub, int* info, int col_base) {
    extern __shared__ __align__(16) double dsm[];
    double* T = dsm;                       // [128][DIAG_DS], row-major: T[r*DS + c]
    double* W = dsm + 128 * DIAG_DS;       // [3][32][33] scratch (S_j blocks of the inverse)
    double* rdiag = W + 96 * 33;           // 1 / L[k][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    DIAG_MARK(0);
    for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {
        const int r = idx & 127, c = idx >> 7;
        T[r * DIAG_DS + c] = (r >= c) ? A[r + (int64_t)c * lda] : 0.0;
    }
    __syncthreads();
    DIAG_MARK(1);

    for (int s = 0; s < 4; ++s) {
        const int c0 = 32 * s;
        if (warp == 0) {
            // (a) pivot block: lane r owns row c0 + r
            double row[32];
            double* Tr = T + (c0 + lane) * DIAG_DS + c0;
#pragma unroll
            for (int k = 0; k < 32; ++k) row[k] = Tr[k];
            // A zero, negative or NaN pivot (nalgebra: is_zero / try_sqrt fails) takes the substitute or records the column;
            // branch-free so the unrolled chain stays one basic block. (Pivots below 1e-300 count as zero: the MUFU seed
            // flushes subnormals.)
            const bool sub_ok = has_sub && sub > 0.0;
            int badcol = 1 << 30;
            auto checked = [&](double d, int k) -> double {
                const bool ok = d > 1e-300;
                if (!ok && !sub_ok) badcol = min(badcol, k);
                return ok ? d : (sub_ok ? sub : nan(""));
            };
            double dk = checked(__shfl_sync(0xffffffffu, row[0], 0), 0);
            double rs = rsqrt_pos(dk);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                // sqrt(dk) = dk * rs with one correction step (<= 1 ulp); only rs is on the critical path
                double sk = dk * rs;
                sk = fma(0.5 * rs, fma(-sk, sk, dk), sk);
                row[k] = (lane == k) ? sk : row[k] * rs;  // lanes < k hold don't-care values above the diagonal
                if (lane == k) rdiag[c0 + k] = rs;
                if (lane >= k) Tr[k] = row[k];
                __syncwarp();
                const int kend = k | 7;  // last column of this group of 8
                if (k < kend) {
#pragma unroll
                    for (int c = k + 1; c <= kend; ++c) row[c] = fma(-row[k], T[(c0 + c) * DIAG_DS + c0 + k], row[c]);
                } else if (k < 31) {
#pragma unroll
                    for (int c = k + 1; c < 32; ++c)
#pragma unroll
                        for (int kk = k - 7; kk <= k; ++kk)
                            row[c] = fma(-row[kk], T[(c0 + c) * DIAG_DS + c0 + kk], row[c]);
                }
                if (k < 31) {
                    dk = checked(__shfl_sync(0xffffffffu, row[k + 1], k + 1), k + 1);
                    rs = rsqrt_pos(dk);
                }
            }
            if (badcol < 32 && lane == 0) atomicCAS(info, 0, col_base + c0 + badcol + 1);
        } else if (s > 0) {
            diag_inv_rowblock(T, W, rdiag, s - 1, warp - 1, DIAG_THREADS / 32 - 1, lane, 1);
        }
        __syncthreads();
        DIAG_MARK(2 + 3 * s);
        const int tn = 128 - c0 - 32;  // rows (and columns) left below / right of the pivot block
        // (b) rows below the pivot block: x * L_pp^T = a, one thread per row, right-looking so updates are independent
        if (tid < tn) {
            double x[32];
            double* Tr = T + (c0 + 32 + tid) * DIAG_DS + c0;
#pragma unroll
            for (int k = 0; k < 32; ++k) x[k] = Tr[k];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                x[k] *= rdiag[c0 + k];
#pragma unroll
                for (int c = k + 1; c < 32; ++c) x[c] = fma(-x[k], T[(c0 + c) * DIAG_DS + c0 + k], x[c]);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) Tr[k] = x[k];
        }
        __syncthreads();
        DIAG_MARK(3 + 3 * s);
        // (c) trailing update inside the tile: A[r][c] -= sum_k X[r][k] X[c][k]
        if (s == 0) diag_syrk<96>(T, c0, warp, lane);
        else if (s == 1) diag_syrk<64>(T, c0, warp, lane);
        else if (s == 2) diag_syrk<32>(T, c0, warp, lane);
        __syncthreads();
        DIAG_MARK(4 + 3 * s);
    }

    // the factor goes back to the matrix (lower part only)
    for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {
        const int r = idx & 127, c = idx >> 7;
        if (r >= c) A[r + (int64_t)c * lda] = T[r * DIAG_DS + c];
    }
    DIAG_MARK(14);
    diag_inv_rowblock(T, W, rdiag, 3, warp, DIAG_THREADS / 32, lane, 0);  // barrier 0 with all threads == __syncthreads
    __syncthreads();
    DIAG_MARK(18);
    for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {
        const int r = idx & 127, c = idx >> 7;
        inv[r + c * 128] = (r > c) ? T[c * DIAG_DS + r] : (r == c ? rdiag[r] : 0.0);
    }
    for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {  // transposed copy for the adjoint solves
        const int c = idx & 127, r = idx >> 7;
        invT[c + r * 128] = (r > c) ? T[c * DIAG_DS + r] : (r == c ? rdiag[r] : 0.0);
    }
    DIAG_MARK(19);
}

// X = L_jj^-1 of every 128 x 128 diagonal tile of an already factored matrix (fgp_upload_state: a deserialised factor has no
// inverse blocks yet).  One CTA per tile, thread c owns column c of X: forward substitution on e_c with the tile's rows
// broadcast from shared memory.  One-off O(n 128^2) work, not on any hot path.
__global__ void __launch_bounds__(128) diag_inverse_kernel(const double* __restrict__ L, int64_t ld, double* __restrict__ inv,
                                                          double* __restrict__ invT) {
    extern __shared__ __align__(16) double Ls[];  // [128][DIAG_DS] row-major, lower triangle
    const int c = threadIdx.x;
    const double* A = L + (int64_t)blockIdx.x * TILE * (ld + 1);
    for (int idx = c; idx < 128 * 128; idx += 128) {
        const int r = idx & 127, cc = idx >> 7;
        Ls[r * DIAG_DS + cc] = (r >= cc) ? A[r + (int64_t)cc * ld] : 0.0;
    }
    __syncthreads();
    double x[128];
    for (int r = 0; r < 128; ++r) {
        double acc = (r == c) ? 1.0 : 0.0;
        const double* lr = Ls + r * DIAG_DS;
        for (int k = 0; k < r; ++k) acc = fma(-lr[k], (k >= c) ? x[k] : 0.0, acc);
        x[r] = (r >= c) ? acc / lr[r] : 0.0;
    }
    double* X = inv + (int64_t)blockIdx.x * TILE * TILE;
    double* XT = invT + (int64_t)blockIdx.x * TILE * TILE;
    for (int r = 0; r < 128; ++r) {
        X[r + c * 128] = x[r];
        XT[c + r * 128] = x[r];
    }
}

void launch_diag_inverse(const double* L, int64_t ld, int64_t nb, double* inv, double* invT, cudaStream_t st) {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (!done) {
        cudaFuncSetAttribute(diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * DIAG_DS * 8);
        done = true;
    }
    diag_inverse_kernel<<<(unsigned)nb, 128, 128 * DIAG_DS * 8, st>>>(L, ld, inv, invT);
}

cudaError_t potrf_prepare() {
    static bool done_dev[64] = {};
    bool& done = *per_device_flag(done_dev);
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = gemm_nt_prepare();
    if (e == cudaSuccess) done = true;
    return e;
}

// block columns [J, Jend) of one panel, left-looking inside the panel; every launch goes to c.st
void factor_panel(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, double* invdiag, double* invdiagT,
                         int has_sub, double sub, int* info, const LaunchCtx& c, PotrfCounters* cnt, const PanelSide* side,
                         const std::function<void(int64_t)>* column_done) {
    const int64_t nb = np / TILE;
    for (int64_t j = J; j < Jend; ++j) {
        double* Ajj = A + j * TILE + j * TILE * lda;
        {
            ProfScope ps(c, PROF_POTRF_DIAG, 128.0 * 128.0 * 128.0 / 3.0);
            potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM_BYTES, c.st>>>(Ajj, lda, invdiag + j * TILE * TILE,
                                                                 invdiagT + j * TILE * TILE, has_sub, sub, info,
                                                                 (int)(j * TILE));
        }
        cnt->launches += 1;
        if (j + 1 < nb) {
            GemmArgs g{};
            g.C = Ajj + TILE; g.ldc = lda;
            g.A = Ajj + TILE; g.lda = lda;
            g.B = invdiag + j * TILE * TILE; g.ldb = TILE;
            g.M = (int)(np - (j + 1) * TILE); g.N = TILE; g.K = TILE;
            g.alpha = 1.0; g.beta_one = 0; g.lower = 0; g.k_from_tile = 0;
            cnt->launches += gemm_nt_launch(g, c) > 0;
        }
        if (column_done) (*column_done)(j);
        if (j + 1 < Jend) {
            // rank-128 update of the panel's remaining block columns by block column j
            if (side && side->st) {
                LaunchCtx sc = c;
                sc.st = side->st;
                cudaEventRecord(side->ev_fork, c.st);            // block column j is final
                cudaStreamWaitEvent(c.st, side->ev_join, 0);     // the side stream's earlier work on block column j+1 is done
                trailing_update(A, lda, np, j, j + 1, j + 1, j + 2, c, cnt);
                if (j + 2 < Jend) {
                    cudaStreamWaitEvent(side->st, side->ev_fork, 0);
                    trailing_update(A, lda, np, j, j + 1, j + 2, Jend, sc, cnt);
                    cudaEventRecord(side->ev_join, side->st);
                }
            } else {
                trailing_update(A, lda, np, j, j + 1, j + 1, Jend, c, cnt);
            }
        }
    }
}

// trailing block columns [c0, c1) (rows >= c0, trapezoid on/below the diagonal) -= P P^T restricted to them, P = panel [J, Jend)
void trailing_update(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, int64_t c0, int64_t c1,
                            const LaunchCtx& c, PotrfCounters* cnt) {
    if (c0 >= c1) return;
    GemmArgs g{};
    g.C = A + c0 * TILE + c0 * TILE * lda; g.ldc = lda;
    g.A = A + c0 * TILE + J * TILE * lda; g.lda = lda;
    g.B = g.A; g.ldb = lda;
    g.M = (int)(np - c0 * TILE); g.N = (int)((c1 - c0) * TILE); g.K = (int)((Jend - J) * TILE);
    g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.k_from_tile = 0;
    cnt->launches += gemm_nt_launch(g, c) > 0;
}

// block columns [c0, c1) of the trailing matrix, rows >= c0 (the trapezoid whose first diagonal tile is (c0, c0)) -= P P^T,
// P = block columns [J, Jend): same tiles and arithmetic as the corresponding part of trailing_update(.., c_first <= c0, ..)
void trailing_update_cols(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, int64_t c0, int64_t c1,
                          const LaunchCtx& c, PotrfCounters* cnt) {
    trailing_update(A, lda, np, J, Jend, c0, c1, c, cnt);
}

void potrf_lower(double* A, int64_t lda, int64_t np, int64_t jb_begin, double* invdiag, double* invdiagT, int has_sub,
                 double sub, int* info, const LaunchCtx& st, const PotrfLookahead* la, PotrfCounters* cnt,
                 const std::function<void()>* after_first_panel_may_start) {
    const int64_t nb = np / TILE;
    const int64_t PANEL_TILES = panel_tiles(np);
    if (jb_begin >= nb) {
        if (after_first_panel_may_start) (*after_first_panel_may_start)();
        return;
    }
    if (!la || !la->panel) {
        if (after_first_panel_may_start) (*after_first_panel_may_start)();
        // single stream, no look-ahead: panel, then the whole trailing update
        for (int64_t J = jb_begin; J < nb; J += PANEL_TILES) {
            const int64_t Jend = (J + PANEL_TILES < nb) ? J + PANEL_TILES : nb;
            factor_panel(A, lda, np, J, Jend, invdiag, invdiagT, has_sub, sub, info, st, cnt);
            trailing_update(A, lda, np, J, Jend, Jend, nb, st, cnt);
        }
        return;
    }
    // One-panel look-ahead on two streams.  Panel stream P (high priority): update of the NEXT panel's block columns by the
    // current panel, then the factorisation of that next panel.  Main stream M: update of all remaining block columns by
    // the current panel.  P's work for panel J+1 overlaps M's big SYRK for panel J.
    LaunchCtx pc = st;
    pc.st = la->panel;
    cudaEventRecord(la->ev_trail, st.st);  // P starts after whatever precedes the factorisation on M (Gram assembly)
    cudaStreamWaitEvent(la->panel, la->ev_trail, 0);
    int64_t J = jb_begin;
    int64_t Jend = (J + PANEL_TILES < nb) ? J + PANEL_TILES : nb;
    {
        // the first panel: nothing has run on the side stream for it, ev_side carries an old (completed) record at most
        const PanelSide ps{la->side, la->ev_fork, la->ev_side};
        factor_panel(A, lda, np, J, Jend, invdiag, invdiagT, has_sub, sub, info, pc, cnt, la->side ? &ps : nullptr);
    }
    cudaEventRecord(la->ev_panel, la->panel);
    bool first = true;
    if (after_first_panel_may_start) {
        (*after_first_panel_may_start)();       // M: e.g. the Gram columns behind the first panel, while P factors it
        cudaEventRecord(la->ev_trail, st.st);   // ... which P must see before its first look-ahead update
        first = false;
    }
    while (Jend < nb) {
        const int64_t Jend2 = (Jend + PANEL_TILES < nb) ? Jend + PANEL_TILES : nb;
        cudaStreamWaitEvent(st.st, la->ev_panel, 0);                   // M: panel [J, Jend) is final
        if (!first) cudaStreamWaitEvent(la->panel, la->ev_trail, 0);   // P: previous trailing update reached columns >= Jend
        if (la->side && Jend2 - Jend > 1) {
            // look-ahead columns: block column 0 of the next panel on P (its factorisation follows at once), the others on
            // the side stream, joined before P touches block column 1
            LaunchCtx sc = st;
            sc.st = la->side;
            cudaEventRecord(la->ev_side, la->panel);                  // side: whatever P waited for (previous trailing update)
            cudaStreamWaitEvent(la->side, la->ev_side, 0);
            trailing_update(A, lda, np, J, Jend, Jend, Jend + 1, pc, cnt);
            trailing_update_cols(A, lda, np, J, Jend, Jend + 1, Jend2, sc, cnt);
            cudaEventRecord(la->ev_side, la->side);
            const PanelSide ps{la->side, la->ev_fork, la->ev_side};
            factor_panel(A, lda, np, Jend, Jend2, invdiag, invdiagT, has_sub, sub, info, pc, cnt, &ps);
        } else {
            trailing_update(A, lda, np, J, Jend, Jend, Jend2, pc, cnt);    // look-ahead columns
            factor_panel(A, lda, np, Jend, Jend2, invdiag, invdiagT, has_sub, sub, info, pc, cnt);
        }
        cudaEventRecord(la->ev_panel, la->panel);
        trailing_update(A, lda, np, J, Jend, Jend2, nb, st, cnt);      // the rest
        cudaEventRecord(la->ev_trail, st.st);
        first = false;
        J = Jend;
        Jend = Jend2;
    }
    cudaStreamWaitEvent(st.st, la->ev_panel, 0);  // join: everything after the factorisation runs on M
}

// ---------------------------------------------------------------------------------------------------------------------
// head schedule (contract in potrf.cuh)
int64_t potrf_lower_head(double* A, int64_t lda, int64_t np, int64_t jb_begin, const PotrfWork& w, int64_t p0, int has_sub,
                         double sub, int* info, const LaunchCtx& st, const PotrfLookahead* la, PotrfCounters* cnt,
                         const std::function<void()>* after_first_panel_may_start) {
    constexpr int64_t PT = HEAD_PANEL / TILE;
    // below this many rows the whole panel solve is one sub-wave launch on the panel stream (no split into top / rest)
    constexpr int64_t SOLVE_SPLIT_ROWS = 4096;
    const int64_t nb = np / TILE;
    if (jb_begin >= nb) {
        if (after_first_panel_may_start) (*after_first_panel_may_start)();
        return 0;
    }
    const bool two = la && la->panel;
    LaunchCtx pc = st, sc = st;
    if (two) {
        pc.st = la->panel;
        sc.st = la->side ? la->side : st.st;
    }
    const int64_t npanels = (nb - jb_begin + PT - 1) / PT;
    cudaMemsetAsync(w.sync + p0 * HEAD_SYNC_INTS, 0, (size_t)npanels * HEAD_SYNC_INTS * sizeof(int), st.st);
    auto head = [&](int64_t J, int64_t Jend, int64_t p) {
        launch_potrf_head(A + J * TILE * (lda + 1), lda, (int)(Jend - J), w.inv + J * TILE * TILE,
                          w.W + (p0 + p) * HEAD_PANEL * HEAD_PANEL, w.P, w.sync + (p0 + p) * HEAD_SYNC_INTS, has_sub, sub, info,
                          (int)(J * TILE), pc);
        cnt->launches += 1;
    };
    // pbuf rows [r0, r1) = A21[r0:r1, :] W^T
    auto solve = [&](int64_t J, int64_t Jend, int64_t p, int64_t r0, int64_t r1, const LaunchCtx& c) {
        if (r1 <= r0) return;
        const int64_t rows = np - Jend * TILE;
        GemmArgs g{};
        g.C = w.pbuf[p & 1] + r0; g.ldc = rows;
        g.A = A + Jend * TILE + r0 + J * TILE * lda; g.lda = lda;
        g.B = w.W + (p0 + p) * HEAD_PANEL * HEAD_PANEL; g.ldb = HEAD_PANEL;
        g.M = (int)(r1 - r0); g.N = (int)((Jend - J) * TILE); g.K = g.N;
        g.alpha = 1.0; g.beta_one = 0; g.lower = 0; g.k_upto_col = 1;
        cnt->launches += gemm_nt_launch(g, c) > 0;
    };
    // trailing matrix behind the panel (origin (Jend, Jend)) -= pbuf pbuf^T: the first `ntop` tile rows only (skip = 0,
    // ntop > 0: the next panel's diagonal block) or everything but the first `skip` tile rows
    auto update = [&](int64_t J, int64_t Jend, int64_t p, int64_t ntop, int64_t skip, const LaunchCtx& c) {
        const int64_t rows = np - Jend * TILE;
        GemmArgs g{};
        g.C = A + Jend * TILE * (lda + 1); g.ldc = lda;
        g.A = w.pbuf[p & 1]; g.lda = rows;
        g.B = g.A; g.ldb = rows;
        g.M = (int)(ntop > 0 ? ntop * TILE : rows); g.N = g.M; g.K = (int)((Jend - J) * TILE);
        g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.row_skip = (int)skip;
        cnt->launches += gemm_nt_launch(g, c) > 0;
    };
    // the same update (skip > 0: everything but the first `skip` tile rows) on tcgen05 when enough rows are left: digit slices
    // of the whole solved panel, then exact int8 products (csrc/ozaki.cuh)
    auto update_big = [&](int64_t J, int64_t Jend, int64_t p, int64_t skip, const LaunchCtx& c) {
        const int64_t rows = np - Jend * TILE;
        if (!w.oz_digits || rows < OZ_MIN_ROWS) {
            update(J, Jend, p, 0, skip, c);
            return;
        }
        const int K = (int)((Jend - J) * TILE);
        int8_t* dg = w.oz_digits;
        double* sc = w.oz_scale;
        if (w.oz_off_bytes && w.oz_off_bytes[p0 + p] >= 0) {   // kept for the solves of predict (trsm.cuh)
            dg += w.oz_off_bytes[p0 + p];
            sc += w.oz_off_rows[p0 + p];
        }
        ozaki_slice_launch(w.pbuf[p & 1], rows, rows, K, dg, sc, c);
        GemmArgs g{};
        g.C = A + Jend * TILE * (lda + 1); g.ldc = lda;
        g.M = (int)rows; g.N = g.M; g.K = K;
        g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.row_skip = (int)skip;
        cnt->launches += 2 + (ozaki_update_launch(g, dg, sc, dg, sc, 0, c) > 0);
    };
    auto copy_back = [&](int64_t J, int64_t Jend, int64_t p, cudaStream_t s) {
        const int64_t rows = np - Jend * TILE;
        cudaMemcpy2DAsync(A + Jend * TILE + J * TILE * lda, lda * sizeof(double), w.pbuf[p & 1], rows * sizeof(double),
                          rows * sizeof(double), (size_t)((Jend - J) * TILE), cudaMemcpyDeviceToDevice, s);
    };

    int64_t J = jb_begin, Jend = std::min(J + PT, nb), p = 0;
    if (!two) {
        if (after_first_panel_may_start) (*after_first_panel_may_start)();
        for (;; ++p) {
            head(J, Jend, p);
            const int64_t rows = np - Jend * TILE;
            if (rows == 0) break;
            solve(J, Jend, p, 0, rows, st);
            update_big(J, Jend, p, 0, st);
            copy_back(J, Jend, p, st.st);
            J = Jend;
            Jend = std::min(J + PT, nb);
        }
        launch_transpose_tiles(w.inv + jb_begin * TILE * TILE, w.invT + jb_begin * TILE * TILE, nb - jb_begin, st.st);
        cnt->launches += 1;
        return p + 1;
    }
    cudaEventRecord(la->ev_trail, st.st);  // P starts after whatever precedes the factorisation on M (Gram of the first panel)
    cudaStreamWaitEvent(la->panel, la->ev_trail, 0);
    head(J, Jend, 0);
    cudaEventRecord(la->ev_panel, la->panel);
    if (after_first_panel_may_start) (*after_first_panel_may_start)();  // M: Gram columns behind the first panel, while P factors it
    cudaEventRecord(la->ev_trail, st.st);
    bool copies[2] = {false, false};
    for (;; ++p) {
        const int64_t rows = np - Jend * TILE;
        if (rows == 0) break;
        const int64_t Jend2 = std::min(Jend + PT, nb), nt2 = Jend2 - Jend;
        const int64_t top = (rows <= SOLVE_SPLIT_ROWS) ? rows : nt2 * TILE;
        cudaStreamWaitEvent(st.st, la->ev_panel, 0);      // M: head(p) done (W_p ready)
        if (copies[p & 1]) {                              // pbuf[p & 1] has left for L (panel p-2)
            cudaStreamWaitEvent(st.st, w.ev_copy[p & 1], 0);
            cudaStreamWaitEvent(la->panel, w.ev_copy[p & 1], 0);
        }
        cudaStreamWaitEvent(la->panel, la->ev_trail, 0);  // P: the previous trailing update reached this panel's rows below
        solve(J, Jend, p, 0, top, pc);
        cudaEventRecord(w.ev_top, la->panel);
        if (top < rows) {
            solve(J, Jend, p, top, rows, st);
            cudaEventRecord(w.ev_rest, st.st);
        }
        update(J, Jend, p, nt2, 0, pc);                   // the next panel's diagonal block ...
        head(Jend, Jend2, p + 1);                         // ... and its head, while M applies panel p to everything else
        cudaEventRecord(la->ev_panel, la->panel);
        cudaStreamWaitEvent(st.st, w.ev_top, 0);
        update_big(J, Jend, p, nt2, st);
        cudaEventRecord(la->ev_trail, st.st);
        cudaStreamWaitEvent(sc.st, w.ev_top, 0);
        if (top < rows) cudaStreamWaitEvent(sc.st, w.ev_rest, 0);
        copy_back(J, Jend, p, sc.st);
        cudaEventRecord(w.ev_copy[p & 1], sc.st);
        copies[p & 1] = true;
        J = Jend;
        Jend = Jend2;
    }
    cudaStreamWaitEvent(st.st, la->ev_panel, 0);  // join: everything after the factorisation runs on M
    for (int i = 0; i < 2; ++i)
        if (copies[i]) cudaStreamWaitEvent(st.st, w.ev_copy[i], 0);
    launch_transpose_tiles(w.inv + jb_begin * TILE * TILE, w.invT + jb_begin * TILE * TILE, nb - jb_begin, st.st);
    cnt->launches += 1;
    return p + 1;
}

}  // namespace fgp
