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
__global__ void __launch_bounds__(256) col_mean_kernel(const double* __restrict__ src, int64_t ld, int64_t n, double* mu) {
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
__global__ void __launch_bounds__(256)
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
// Forward (L x = b), right-looking over 128-blocks. Launch j = 0 .. nb-1 with grid = nb - j:
//   block 0 solves x_j = inv_j * b_j (b_j is final: every earlier launch has been applied to it);
//   block t > 0 waits for nothing: it applies the PREVIOUS solution, b_R -= L[R, j-1] x_{j-1}, R = j + t ... so the
// kernel is split in two phases per launch: (1) every block R = j + blockIdx.x applies x_{j-1} (when j > 0),
// (2) block 0 solves x_j.
__global__ void __launch_bounds__(128)
trsv_fwd_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ inv, double* b, double* x, int j) {
    __shared__ double xs[128];
    const int r = threadIdx.x;
    const int R = j + blockIdx.x;
    const int64_t row = (int64_t)R * 128 + r;
    double v = b[row];
    if (j > 0) {
        xs[r] = x[(int64_t)(j - 1) * 128 + r];
        __syncthreads();
        const double* Lp = L + row + (int64_t)(j - 1) * 128 * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int c = 0; c < 128; c += 4) {
            a0 = fma(Lp[(int64_t)c * ld], xs[c], a0);
            a1 = fma(Lp[(int64_t)(c + 1) * ld], xs[c + 1], a1);
            a2 = fma(Lp[(int64_t)(c + 2) * ld], xs[c + 2], a2);
            a3 = fma(Lp[(int64_t)(c + 3) * ld], xs[c + 3], a3);
        }
        v -= (a0 + a1) + (a2 + a3);
        b[row] = v;
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        xs[r] = v;
        __syncthreads();
        const double* ip = inv + (int64_t)R * 128 * 128 + r;
        double a0 = 0, a1 = 0;
        for (int c = 0; c < 128; c += 2) {
            a0 = fma(ip[c * 128], xs[c], a0);
            a1 = fma(ip[(c + 1) * 128], xs[c + 1], a1);
        }
        x[row] = a0 + a1;
    }
}

// Adjoint (L^T x = b), right-looking from the last block. Launch j = nb-1 .. 0 with grid = j + 1:
//   (1) when j < nb-1 every block C = blockIdx.x <= j applies the previous solution: b_C -= L[j+1, C]^T x_{j+1};
//   (2) block C == j solves x_j = inv_j^T b_j.
__global__ void __launch_bounds__(128)
trsv_adj_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ invT, double* b, double* x, int j,
                int nb) {
    __shared__ double xs[128];
    __shared__ double bs[128];
    const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
    const int Cb = blockIdx.x;
    bs[r] = b[(int64_t)Cb * 128 + r];
    if (j < nb - 1) {
        xs[r] = x[(int64_t)(j + 1) * 128 + r];
        __syncthreads();
        for (int cc = 0; cc < 32; ++cc) {
            const int cl = warp * 32 + cc;
            const double* Lp = L + (int64_t)(j + 1) * 128 + ((int64_t)Cb * 128 + cl) * ld;
            double p = Lp[lane] * xs[lane] + Lp[lane + 32] * xs[lane + 32] + Lp[lane + 64] * xs[lane + 64] +
                       Lp[lane + 96] * xs[lane + 96];
            p = warp_sum(p);
            if (lane == 0) bs[cl] -= p;
        }
        __syncthreads();
        b[(int64_t)Cb * 128 + r] = bs[r];
    }
    __syncthreads();
    if (Cb == j) {
        const double* ip = invT + (int64_t)Cb * 128 * 128 + r;
        double a0 = 0, a1 = 0;
        for (int c = 0; c < 128; c += 2) {
            a0 = fma(ip[c * 128], bs[c], a0);
            a1 = fma(ip[(c + 1) * 128], bs[c + 1], a1);
        }
        x[(int64_t)Cb * 128 + r] = a0 + a1;
    }
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
__global__ void __launch_bounds__(128)
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

__global__ void fill_kernel(double* p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace fgp
