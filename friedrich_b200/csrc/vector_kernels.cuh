// vector_kernels.cuh — the O(n d) / O(n^2) memory-bound helpers around the tensor-pipe kernels:
// input conversion (column-major host layout -> padded row-major device layout, centring, norms), blocked
// single-right-hand-side triangular solves (alpha = K^-1 y; first half = `ol` of likelihood, mod.rs:203),
// row reductions of the transposed solve buffer (posterior mean mod.rs:241 / variance diagonal mod.rs:266-270),
// and the likelihood terms (mod.rs:196-220).
#pragma once

#include "common.cuh"
#include "kernel_eval.cuh"

namespace fgp {

// mu[k] = mean_i src[i + k*ld], i < n.  One block per column, deterministic tree.
static __global__ void __launch_bounds__(256) col_mean_kernel(const double* __restrict__ src, int64_t ld, int64_t n, double* mu) {
    __shared__ double red[256];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += src[i + k * ld];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) mu[k] = red[0] / (double)n;
}

// src: n x d column-major (ld). Writes rows [row0, row0 + rows_total) of the row-major [.][dp] arrays:
// raw, centred (raw - mu for k < d), and the squared norms of both; rows >= row0 + n and columns >= d are zero.
static __global__ void __launch_bounds__(256)
convert_points_kernel(const double* __restrict__ src, int64_t ld, int64_t n, int d, int dp, const double* __restrict__ mu,
                      int64_t row0, int64_t rows_total, double* xr, double* xc, double* nrm_c, double* nrm_r) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows_total) return;
    double* pr = xr + (row0 + i) * dp;
    double* pc = xc + (row0 + i) * dp;
    double sc = 0.0, sr = 0.0;
    for (int k = 0; k < dp; ++k) {
        double v = 0.0, c = 0.0;
        if (i < n && k < d) {
            v = src[i + (int64_t)k * ld];
            c = v - mu[k];
        }
        pr[k] = v;
        pc[k] = c;
        sc = fma(c, c, sc);
        sr = fma(v, v, sr);
    }
    nrm_c[row0 + i] = sc;
    nrm_r[row0 + i] = sr;
}

// ---- blocked single-RHS solves with the inverted diagonal blocks ------------------------------------------------
// Both sweeps are HBM-bound (they read L once: 8 n^2/2 bytes) but latency-critical: nb dependent block steps.  Every
// CTA therefore issues ALL its loads of a 128x128 tile up front (512 threads x 32 independent 8-byte loads), reduces
// through shared memory in a fixed order (deterministic), and the step count is one launch per block.
constexpr int TRSV_THREADS = 512;

// y[r] = sum_c M[r + c*ldm] * xs[c], r < 128, c < 128; thread (r = tid & 127, quarter = tid >> 7) sums 32 columns with 4
// interleaved accumulators, the quarters are added in order.  Result valid in threads with tid < 128.  Split in two so the
// tile's loads can be issued BEFORE the vector is known (the wavefront solves below):
__device__ __forceinline__ void tile_load_512(const double* __restrict__ M, int64_t ldm, double (&v)[32]) {
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    const double* p = M + r + (int64_t)(32 * h) * ldm;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = __ldcg(p + (int64_t)c * ldm);
}
__device__ __forceinline__ double tile_apply_512(const double (&v)[32], const double* xs, double* red) {
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        a0 = fma(v[c], xs[32 * h + c], a0);
        a1 = fma(v[c + 1], xs[32 * h + c + 1], a1);
        a2 = fma(v[c + 2], xs[32 * h + c + 2], a2);
        a3 = fma(v[c + 3], xs[32 * h + c + 3], a3);
    }
    red[h * 128 + r] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    double out = 0.0;
    if (threadIdx.x < 128) out = (red[r] + red[128 + r]) + (red[256 + r] + red[384 + r]);
    __syncthreads();
    return out;
}
// the same product with the 128 x 128 tile resident in shared memory (column-major, ld = 128)
__device__ __forceinline__ double tile_apply_smem_512(const double* Ms, const double* xs, double* red) {
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    const double* p = Ms + r + 32 * h * 128;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        a0 = fma(p[c * 128], xs[32 * h + c], a0);
        a1 = fma(p[(c + 1) * 128], xs[32 * h + c + 1], a1);
        a2 = fma(p[(c + 2) * 128], xs[32 * h + c + 2], a2);
        a3 = fma(p[(c + 3) * 128], xs[32 * h + c + 3], a3);
    }
    red[h * 128 + r] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    double out = 0.0;
    if (threadIdx.x < 128) out = (red[r] + red[128 + r]) + (red[256 + r] + red[384 + r]);
    __syncthreads();
    return out;
}

// ---- wavefront triangular solves: ONE launch per solve instead of one per block column -------------------------------
// Block i of the grid owns the 128 unknowns of block row i. It consumes the solution blocks it depends on as their owners
// publish them (a flag per block in global memory, release/acquire), so the 128-step dependency chain costs a flag
// round trip per step instead of a kernel launch.  A CTA takes its block index from an atomic TICKET at kernel start, not
// from blockIdx: a block only ever waits for lower tickets, i.e. for CTAs that are already running, so the wait cannot
// deadlock however many blocks are resident and in whatever order the hardware dispatches them.  Everything that does not depend on
// the awaited block is fetched BEFORE the wait: the block's inverse diagonal tile arrives in shared memory by TMA bulk
// copies at kernel start, and the L tile of each step is loaded into registers ahead of its flag.  The arithmetic (order
// of the block updates j = 0, 1, ... and the mat-vec reductions) does not depend on the timing: results are deterministic.
constexpr int TRSV_WAVE_SMEM = 128 * 128 * 8 + 16;
__device__ __forceinline__ void wave_wait(const int* flag) {
    if (threadIdx.x == 0) {
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        } while (v == 0);
    }
    __syncthreads();
}
__device__ __forceinline__ void wave_publish(int* flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
// 128 columns x 1 KiB of a contiguous 128 x 128 tile -> shared memory, completion on `bar` (lanes 0..127 of the CTA)
__device__ __forceinline__ void wave_fetch_tile(double* dst, const double* src, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, 128 * 128 * 8);
    __syncthreads();
    if (threadIdx.x < 128) tma_load_1d(dst + threadIdx.x * 128, src + threadIdx.x * 128, 1024, bar);
}

// Forward: L x = b.  x_i = inv_i (b_i - sum_{j<i} L[i,j] x_j)
static __global__ void __launch_bounds__(TRSV_THREADS)
trsv_fwd_wave_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ inv, const double* b, double* x,
                     int* flags, int* ticket) {  // b may alias x: a block reads its segment of b before it writes that segment of x
    extern __shared__ __align__(128) unsigned char wave_smem[];
    double* inv_s = reinterpret_cast<double*>(wave_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wave_smem + 128 * 128 * 8);
    __shared__ double xs[128];
    __shared__ double red[512];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int r = threadIdx.x & 127;
    const int i = s_ticket;
    const int64_t row0 = (int64_t)i * 128;
    wave_fetch_tile(inv_s, inv + (int64_t)i * 128 * 128, bar);
    double v = 0.0;
    if (threadIdx.x < 128) v = b[row0 + r];
    double t[32];
    for (int j = 0; j < i; ++j) {
        tile_load_512(L + row0 + (int64_t)j * 128 * ld, ld, t);  // independent of x_j: in flight while we wait for it
        wave_wait(flags + j);
        if (threadIdx.x < 128) xs[r] = __ldcg(x + (int64_t)j * 128 + r);
        __syncthreads();
        v -= tile_apply_512(t, xs, red);
    }
    if (threadIdx.x < 128) xs[r] = v;
    mbar_wait(bar, 0);
    __syncthreads();
    const double s = tile_apply_smem_512(inv_s, xs, red);
    if (threadIdx.x < 128) x[row0 + r] = s;
    wave_publish(flags + i);
}

// Forward solve for up to 16 right-hand sides AT ONCE (the latency path of predict for a handful of queries: one launch
// and one walk over L instead of one per query).  X is np x q (one right-hand side per column, ld = ldx), solved in place.
// Same ticket / flag protocol as above; a block publishes after all its q segments are written.  Thread (r, h) = (tid & 127,
// tid >> 7) holds row r, columns 32 h .. 32 h + 31 of the current L tile in registers (loaded once per step, before the
// wait) and applies it to every right-hand side; the x values are warp-uniform shared-memory broadcasts, the four column
// quarters meet in shared memory.  (Per step and right-hand side the SM has 128 x 128 FMAs to do: 256 cycles at its fp64 rate.)
constexpr int TRSV_MULTI_QMAX = 16;
// One 128 x 128 tile (thread (r, h): row r, columns 32 h .. 32 h + 31 in v) applied to q right-hand sides in shared memory
// (xs[j][128]); the four column quarters' partial sums go to red[j][h][r].  The x values are warp-uniform 16-byte broadcasts
// (two per load: the loop is bound by shared-memory loads, not by the fp64 pipe); four accumulators per right-hand side as in
// tile_apply_512.
__device__ __forceinline__ void trsv_multi_apply(const double (&v)[32], const double* xs, double* red, int q, int r, int h) {
#pragma unroll
    for (int j = 0; j < TRSV_MULTI_QMAX; ++j) {
        if (j < q) {
            const double2* xj = reinterpret_cast<const double2*>(xs + j * 128 + 32 * h);
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const double2 x01 = xj[c >> 1], x23 = xj[(c >> 1) + 1];
                a0 = fma(v[c], x01.x, a0);
                a1 = fma(v[c + 1], x01.y, a1);
                a2 = fma(v[c + 2], x23.x, a2);
                a3 = fma(v[c + 3], x23.y, a3);
            }
            red[(j * 4 + h) * 128 + r] = (a0 + a1) + (a2 + a3);
        }
    }
}
constexpr int TRSV_MULTI_SMEM = TRSV_WAVE_SMEM + (TRSV_MULTI_QMAX * 128 + TRSV_MULTI_QMAX * 512) * 8;
static __global__ void __launch_bounds__(TRSV_THREADS)
trsv_fwd_wave_multi_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ inv, double* X, int64_t ldx, int q,
                           int* flags, int* ticket) {
    extern __shared__ __align__(128) unsigned char wave_smem[];
    double* inv_s = reinterpret_cast<double*>(wave_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wave_smem + 128 * 128 * 8);
    double* xs = reinterpret_cast<double*>(wave_smem + TRSV_WAVE_SMEM);  // [QMAX][128]
    double* red = xs + TRSV_MULTI_QMAX * 128;                            // [QMAX][4][128]
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int tid = threadIdx.x, r = tid & 127, h = tid >> 7;
    const int i = s_ticket;
    const int64_t row0 = (int64_t)i * 128;
    wave_fetch_tile(inv_s, inv + (int64_t)i * 128 * 128, bar);
    double acc[TRSV_MULTI_QMAX];  // threads 0..127: the running right-hand sides of row r
#pragma unroll
    for (int j = 0; j < TRSV_MULTI_QMAX; ++j) acc[j] = (j < q && h == 0) ? X[row0 + r + (int64_t)j * ldx] : 0.0;
    for (int jb = 0; jb < i; ++jb) {
        double v[32];  // independent of x_jb: in flight while we wait for it
        tile_load_512(L + row0 + (int64_t)jb * 128 * ld, ld, v);
        wave_wait(flags + jb);
        for (int idx = tid; idx < q * 128; idx += TRSV_THREADS)
            xs[idx] = __ldcg(X + (int64_t)jb * 128 + (idx & 127) + (int64_t)(idx >> 7) * ldx);
        __syncthreads();
        trsv_multi_apply(v, xs, red, q, r, h);
        __syncthreads();
        if (h == 0) {
#pragma unroll
            for (int j = 0; j < TRSV_MULTI_QMAX; ++j)
                if (j < q) {
                    const double* rj = red + j * 512 + r;
                    acc[j] -= (rj[0] + rj[128]) + (rj[256] + rj[384]);
                }
        }
        // (the next step's xs / red writes come after its own wave_wait barrier: no extra barrier needed here)
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TRSV_MULTI_QMAX; ++j)
        if (j < q && h == 0) xs[j * 128 + r] = acc[j];
    mbar_wait(bar, 0);
    __syncthreads();
    // x_i = inv_i (...) for ALL right-hand sides in one pass: the tile's values once into registers, one barrier pair for the
    // lot (a pass per right-hand side put 2 q barriers on every link of the wavefront chain); per right-hand side the same
    // sums in the same order as tile_apply_smem_512
    {
        const double* p = inv_s + r + 32 * h * 128;
        double w[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) w[c] = p[c * 128];
        trsv_multi_apply(w, xs, red, q, r, h);
        __syncthreads();
        if (h == 0) {
#pragma unroll
            for (int j = 0; j < TRSV_MULTI_QMAX; ++j)
                if (j < q) {
                    const double* rj = red + j * 512 + r;
                    X[row0 + r + (int64_t)j * ldx] = (rj[0] + rj[128]) + (rj[256] + rj[384]);
                }
        }
    }
    wave_publish(flags + i);
}

// Adjoint: L^T x = b, from the last block.  The CTA with ticket g owns block column i = nb-1-g:
// x_i = inv_i^T (b_i - sum_{j>i} L[j,i]^T x_j); warp w owns columns 8w .. 8w+7 of a tile, a lane reads rows lane, lane+32,
// lane+64, lane+96 of each (32 independent loads), then eight shuffle reductions.
static __global__ void __launch_bounds__(TRSV_THREADS)
trsv_adj_wave_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ invT, const double* __restrict__ b,
                     double* x, int* flags, int* ticket, int nb) {
    extern __shared__ __align__(128) unsigned char wave_smem[];
    double* inv_s = reinterpret_cast<double*>(wave_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(wave_smem + 128 * 128 * 8);
    __shared__ double xs[128];
    __shared__ double bs[128];
    __shared__ double red[512];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = nb - 1 - s_ticket;
    wave_fetch_tile(inv_s, invT + (int64_t)i * 128 * 128, bar);
    if (tid < 128) bs[tid] = b[(int64_t)i * 128 + tid];
    for (int j = nb - 1; j > i; --j) {
        const double* Lp = L + (int64_t)j * 128 + ((int64_t)i * 128 + 8 * warp) * ld + lane;
        double v[8][4];  // independent of x_j: in flight while we wait for it
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[c][k] = __ldcg(Lp + (int64_t)c * ld + 32 * k);
        wave_wait(flags + (nb - 1 - j));
        if (tid < 128) xs[tid] = __ldcg(x + (int64_t)j * 128 + tid);
        __syncthreads();
        const double x0 = xs[lane], x1 = xs[lane + 32], x2 = xs[lane + 64], x3 = xs[lane + 96];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double p = (v[c][0] * x0 + v[c][1] * x1) + (v[c][2] * x2 + v[c][3] * x3);
            p = warp_sum(p);
            if (lane == 0) bs[8 * warp + c] -= p;
        }
    }
    mbar_wait(bar, 0);
    __syncthreads();
    const double s = tile_apply_smem_512(inv_s, bs, red);
    if (tid < 128) x[(int64_t)i * 128 + tid] = s;
    wave_publish(flags + (nb - 1 - i));
}

// ---- row reductions of the transposed buffer Bt (qp x np, column-major, ld) ---------------------------------------
// partial[chunk][c] = sum_{i in chunk} f(Bt[c,i]) ; MODE 0: Bt[c,i] * vec[i]   MODE 1: Bt[c,i]^2
constexpr int ROWRED_CHUNK = 128;
template <int MODE>
__global__ void __launch_bounds__(128)
rowreduce_partial_kernel(const double* __restrict__ Bt, int64_t ld, const double* __restrict__ vec, double* partial,
                         int64_t qp) {
    const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * ROWRED_CHUNK;
    const double* p = Bt + c + i0 * ld;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int i = 0; i < ROWRED_CHUNK; i += 4) {
        const double v0 = p[(int64_t)i * ld], v1 = p[(int64_t)(i + 1) * ld], v2 = p[(int64_t)(i + 2) * ld],
                     v3 = p[(int64_t)(i + 3) * ld];
        if (MODE == 0) {
            a0 = fma(v0, vec[i0 + i], a0); a1 = fma(v1, vec[i0 + i + 1], a1);
            a2 = fma(v2, vec[i0 + i + 2], a2); a3 = fma(v3, vec[i0 + i + 3], a3);
        } else {
            a0 = fma(v0, v0, a0); a1 = fma(v1, v1, a1); a2 = fma(v2, v2, a2); a3 = fma(v3, v3, a3);
        }
    }
    partial[(int64_t)blockIdx.y * qp + c] = (a0 + a1) + (a2 + a3);
}

// out[c] = sum_chunks partial ;  with_prior_var: out[c] = k(q_c, q_c) - sum   (mod.rs:266-270)
static __global__ void __launch_bounds__(128)
rowreduce_final_kernel(const double* __restrict__ partial, int chunks, int64_t qp, int64_t q, int variance, DevKernel k,
                       const double* __restrict__ qnorm_raw, double* out) {
    const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (c >= q) return;
    double s = 0.0;
    for (int ch = 0; ch < chunks; ++ch) s += partial[(int64_t)ch * qp + c];
    out[c] = variance ? (kernel_value<KIND_GENERIC>(k, qnorm_raw[c], 0.0) - s) : s;
}

// single-block deterministic reductions -------------------------------------------------------------------------
// MODE 0: sum v[i]^2 ; MODE 1: sum ln|k(x_i,x_i) + noise2| (needs raw norms) ; MODE 2: sum v[i]*w[i]
template <int MODE>
__global__ void __launch_bounds__(256)
reduce_kernel(const double* __restrict__ v, const double* __restrict__ w, int64_t n, DevKernel k, double noise2, double* out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        if (MODE == 0) s = fma(v[i], v[i], s);
        if (MODE == 1) s += log(fabs(kernel_value<KIND_GENERIC>(k, v[i], 0.0) + noise2));
        if (MODE == 2) s = fma(v[i], w[i], s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// one block per column c of a column-major matrix: out[c] = sum_r M[r + c*ld] * vec[r] (MODE 0) / sum_r M[r + c*ld]^2 (MODE 1);
// fixed-order tree reduction (deterministic). The single-query latency path of predict.
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(const double* __restrict__ M, int64_t ld, const double* __restrict__ vec, int64_t n, double* out) {
    __shared__ double red[256];
    const double* p = M + (int64_t)blockIdx.x * ld;
    double a0 = 0.0, a1 = 0.0;
    int64_t i = threadIdx.x;
    for (; i + 256 < n; i += 512) {
        const double v0 = p[i], v1 = p[i + 256];
        if (MODE == 0) { a0 = fma(v0, vec[i], a0); a1 = fma(v1, vec[i + 256], a1); }
        else { a0 = fma(v0, v0, a0); a1 = fma(v1, v1, a1); }
    }
    if (i < n) a0 = (MODE == 0) ? fma(p[i], vec[i], a0) : fma(p[i], p[i], a0);
    red[threadIdx.x] = a0 + a1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

static __global__ void fill_kernel(double* p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace fgp
