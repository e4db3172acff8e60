// lml.cu — LML-gradient reductions and the bandwidth heuristic on the pair-tile engine (contract in lml.cuh).
#include "lml.cuh"

namespace fgp {

template <int KIND, int PMAX>
struct LmlEpi {
    static constexpr int NV = 2 * PMAX + 1;
    DevKernel k;
    const double* kinv;
    int64_t ld;
    const double* alpha;
    int64_t n;
    double* partial;
    double acc[NV];  // [0,P): sum w Kinv G_p   [PMAX, PMAX+P): sum w a_r a_c G_p   [2 PMAX]: tr Kinv
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, double dot, double d2) {
        if (r >= n || c >= n) return;
        const double ki = kinv[r + c * ld];
        const double aa = alpha[r] * alpha[c];
        const double w = (r == c) ? 1.0 : 2.0;  // the gradient matrices are symmetric (algebra/mod.rs:148-149)
        if (r == c) acc[2 * PMAX] += ki;
        if (KIND == KIND_SQEXP || KIND == KIND_MATERN2) {
            double g[2];
            leaf_grad(KIND == KIND_SQEXP ? FGP_K_SQUARED_EXP : FGP_K_MATERN2, k.param, dot, d2, g);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                acc[p] = fma(w * ki, g[p], acc[p]);
                acc[PMAX + p] = fma(w * aa, g[p], acc[PMAX + p]);
            }
        } else {
            double g[FGP_MAX_PARAMS];
            double val;
            const int P = kernel_value_grad(k, dot, d2, &val, g);
            for (int p = 0; p < P; ++p) {
                acc[p] = fma(w * ki, g[p], acc[p]);
                acc[PMAX + p] = fma(w * aa, g[p], acc[PMAX + p]);
            }
        }
    }
    __device__ __forceinline__ void skipped(int64_t, int64_t) {}
    template <class E>
    __device__ __forceinline__ void bind(const E*) {}
    __device__ __forceinline__ void finish(int64_t, int64_t, bool) {
        block_sum_store<NV>(acc, partial + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * NV);
    }
};

// sum of Euclidean distances over the strict lower triangle (fit_bandwidth_mean, kernel.rs:94-113)
struct DistEpi {
    int64_t n;
    double* partial;
    double acc[1];
    __device__ __forceinline__ void operator()(int64_t r, int64_t c, double, double d2) {
        if (r < n && c < r) acc[0] += sqrt(d2);
    }
    __device__ __forceinline__ void skipped(int64_t, int64_t) {}
    template <class E>
    __device__ __forceinline__ void bind(const E*) {}
    __device__ __forceinline__ void finish(int64_t, int64_t, bool) { block_sum_store<1>(acc, partial + (size_t)(blockIdx.y * gridDim.x + blockIdx.x)); }
};

// out[v] = sum_b partial[b*nv + v], b ascending; one thread per value, 4 interleaved partial sums
__global__ void __launch_bounds__(64) partials_final_kernel(const double* __restrict__ partial, int64_t blocks, int nv, double* out) {
    const int v = threadIdx.x;
    if (v >= nv) return;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int64_t b = 0;
    for (; b + 3 < blocks; b += 4) {
        s0 += partial[b * nv + v];
        s1 += partial[(b + 1) * nv + v];
        s2 += partial[(b + 2) * nv + v];
        s3 += partial[(b + 3) * nv + v];
    }
    for (; b < blocks; ++b) s0 += partial[b * nv + v];
    out[v] = (s0 + s1) + (s2 + s3);
}

__global__ void set_identity_kernel(double* A, int64_t ld, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i + i * ld] = 1.0;
}

template <int KIND, int PMAX, int MODE>
void launch_lml(fgp_model* m, const PairArgs& pa, const DevKernel& dk, const double* kinv, int64_t ld, double* partial) {
    LmlEpi<KIND, PMAX> e{};
    e.k = dk;
    e.kinv = kinv;
    e.ld = ld;
    e.alpha = m->alpha.p;
    e.n = m->n;
    e.partial = partial;
    for (int i = 0; i < LmlEpi<KIND, PMAX>::NV; ++i) e.acc[i] = 0.0;
    pair_tile_kernel<MODE, LmlEpi<KIND, PMAX>><<<pair_grid(pa), 256, 0, m->st>>>(pa, e);
    m->launches += 1;
}

int lml_gradient_device(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int scaled,
                               double* scale_out, double* grads) {
    const int64_t np = m->np, nb = np / TILE;
    const int P = kt.nparams;
    // U = L^-T  (upper triangular, column-major np x np)
    CU(m, m->U.reserve((size_t)np * np));
    CU(m, m->Kinv.reserve((size_t)np * np));
    CU(m, cudaMemsetAsync(m->U.p, 0, (size_t)np * np * sizeof(double), m->st));
    set_identity_kernel<<<(unsigned)((np + 255) / 256), 256, 0, m->st>>>(m->U.p, np, np);
    m->launches += 1;
    m->launches += trsm_fwd_t(m->U.p, np, np, m->L.p, m->cap, m->inv.p, 0, nb, nullptr, m->ctx(), true);
    // K^-1 = U U^T, lower triangle
    {
        GemmArgs g{};
        g.C = m->Kinv.p; g.ldc = np;
        g.A = m->U.p; g.lda = np;
        g.B = m->U.p; g.ldb = np;
        g.M = g.N = (int)np; g.K = (int)np;
        g.alpha = 1.0; g.beta_one = 0; g.lower = 1; g.k_from_tile = 1;
        m->launches += gemm_nt_launch(g, m->ctx()) > 0;
    }
    // fused gradient reductions
    PairArgs pa{};
    pa.xa_c = pa.xb_c = m->xc.p;
    pa.xa_r = pa.xb_r = m->xr.p;
    pa.na = pa.nb = m->nc.p;
    pa.dp = (int)m->dp;
    pa.rows = pa.cols = np;
    pa.row_tile0 = 0;
    pa.symmetric = 1;
    const dim3 grid = pair_grid(pa);
    const int64_t blocks = (int64_t)grid.x * grid.y;
    const DevKernel dk = to_dev(kd);
    const int mode = (kt.need_d2 ? PAIR_D2 : 0) | (kt.need_dot ? PAIR_DOT : 0);
    int nv, pmax;
    if (kt.kind == KIND_SQEXP || kt.kind == KIND_MATERN2) {
        pmax = 2; nv = 5;
        CU(m, m->lml_partial.reserve((size_t)blocks * nv));
        if (kt.kind == KIND_SQEXP) launch_lml<KIND_SQEXP, 2, PAIR_D2>(m, pa, dk, m->Kinv.p, np, m->lml_partial.p);
        else launch_lml<KIND_MATERN2, 2, PAIR_D2>(m, pa, dk, m->Kinv.p, np, m->lml_partial.p);
    } else {
        pmax = FGP_MAX_PARAMS; nv = 2 * FGP_MAX_PARAMS + 1;
        CU(m, m->lml_partial.reserve((size_t)blocks * nv));
        if (mode == PAIR_D2) launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_D2>(m, pa, dk, m->Kinv.p, np, m->lml_partial.p);
        else if (mode == PAIR_DOT) launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_DOT>(m, pa, dk, m->Kinv.p, np, m->lml_partial.p);
        else launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_BOTH>(m, pa, dk, m->Kinv.p, np, m->lml_partial.p);
    }
    CU(m, m->scalars.reserve(64));
    partials_final_kernel<<<1, 64, 0, m->st>>>(m->lml_partial.p, blocks, nv, m->scalars.p);
    reduce_kernel<2><<<1, 256, 0, m->st>>>(m->alpha.p, m->alpha.p, m->n, dk, 0.0, m->scalars.p + 50);  // alpha.alpha
    reduce_kernel<2><<<1, 256, 0, m->st>>>(m->y.p, m->alpha.p, m->n, dk, 0.0, m->scalars.p + 51);      // y.alpha
    m->launches += 3;
    FGP_TRY(ensure_pinned(m, 64));
    CU(m, cudaMemcpyAsync(m->pinned, m->scalars.p, 52 * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    const double* h = m->pinned;
    const double scale = scaled ? h[51] / (double)m->n : 1.0;  // optimizer.rs:174
    for (int p = 0; p < P; ++p) {
        double data_fit = h[pmax + p];
        if (scaled) data_fit /= scale;                          // optimizer.rs:186
        grads[p] = (data_fit - h[p]) / 2.0;                     // optimizer.rs:192 / :49
    }
    if (!scaled) grads[P] = noise * (h[50] - h[2 * pmax]);      // optimizer.rs:54-57
    if (scale_out) *scale_out = scale;
    m->kinv_valid = true;
    return FGP_OK;
}

int mean_pair_distance_device(fgp_model* m, double* out) {
    PairArgs pa{};
    pa.xa_c = pa.xb_c = m->xc.p;
    pa.xa_r = pa.xb_r = m->xr.p;
    pa.na = pa.nb = m->nc.p;
    pa.dp = (int)m->dp;
    pa.rows = pa.cols = m->np;
    pa.row_tile0 = 0;
    pa.symmetric = 1;
    const dim3 grid = pair_grid(pa);
    const int64_t blocks = (int64_t)grid.x * grid.y;
    CU(m, m->lml_partial.reserve((size_t)blocks));
    CU(m, m->scalars.reserve(64));
    DistEpi e{m->n, m->lml_partial.p, {0.0}};
    pair_tile_kernel<PAIR_D2, DistEpi><<<grid, 256, 0, m->st>>>(pa, e);
    partials_final_kernel<<<1, 64, 0, m->st>>>(m->lml_partial.p, blocks, 1, m->scalars.p);
    m->launches += 2;
    FGP_TRY(ensure_pinned(m, 64));
    CU(m, cudaMemcpyAsync(m->pinned, m->scalars.p, sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    const double n = (double)m->n;
    *out = m->pinned[0] / ((n * n - n) / 2.0);  // kernel.rs:111-112
    return FGP_OK;
}

void launch_set_identity(double* A, int64_t ld, int64_t n, cudaStream_t st) {
    set_identity_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A, ld, n);
}

}  // namespace fgp
