// lml.cu — LML-gradient reductions and the bandwidth heuristic on the pair-tile engine (contract in lml.cuh).
#include "lml.cuh"

#include <cmath>
#include <vector>
#include "ozaki.cuh"
#include "potrf.cuh"
#include "trsm.cuh"

#include "potrf.cuh"
#include "sharded.cuh"

namespace fgp {

template <int KIND, int PMAX>
struct LmlEpi {
    static constexpr int NV = 2 * PMAX + 1;
    static constexpr bool HAS_FAST = false;
    __device__ __forceinline__ bool fast_ok(int64_t, int64_t, int) const { return false; }
    __device__ __forceinline__ double* fast_base(int64_t, int64_t) const { return nullptr; }
    __device__ __forceinline__ int fast_ld() const { return 0; }
    __device__ __forceinline__ double fast_scale() const { return 1.0; }
    __device__ __forceinline__ double fast_value(double, const double*) const { return 0.0; }
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
    static constexpr bool HAS_FAST = false;
    __device__ __forceinline__ bool fast_ok(int64_t, int64_t, int) const { return false; }
    __device__ __forceinline__ double* fast_base(int64_t, int64_t) const { return nullptr; }
    __device__ __forceinline__ int fast_ld() const { return 0; }
    __device__ __forceinline__ double fast_scale() const { return 1.0; }
    __device__ __forceinline__ double fast_value(double, const double*) const { return 0.0; }
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
    const bool panels = m->w_valid && !m->pstart.empty() && np <= m->cap && m->pbuf[0].p && m->pbuf[1].p;
    const bool tc = m->tcgen05 && np >= OZ_MIN_ROWS && np <= 32768 && ozaki_prepare() == cudaSuccess;
    if (panels) {
        // over the panels of the head schedule (W_p = L11^-1 per panel): two launches per 512 columns; where the fit kept the
        // panel's digit slices the K = 512 updates run on tcgen05 (trsm.cuh)
        OzPanelStore ozs{};
        const bool use_oz = tc && m->ozL_valid && m->ozOffBytes.size() == m->pstart.size() && m->ozDigits.p;
        if (use_oz)
            ozs = OzPanelStore{reinterpret_cast<const int8_t*>(m->ozL.p), m->ozLscale.p, m->ozOffBytes.data(), m->ozOffRows.data(),
                               reinterpret_cast<int8_t*>(m->ozDigits.p), m->ozScale.p};
        m->launches += trsm_fwd_t_panels(m->U.p, np, np, m->L.p, m->cap, m->Wp.p, m->pstart.data(), (int64_t)m->pstart.size(), nb, nb,
                                         m->pbuf[0].p, m->pbuf[1].p, m->ctx(), m->lookahead ? m->st2 : nullptr, m->evA, m->evB, m->evC,
                                         use_oz ? &ozs : nullptr, true);
    } else {
        m->launches += trsm_fwd_t(m->U.p, np, np, m->L.p, m->cap, m->inv.p, 0, nb, nullptr, m->ctx(), true);
    }
    // K^-1 = U U^T, lower triangle
    {
        GemmArgs g{};
        g.C = m->Kinv.p; g.ldc = np;
        g.A = m->U.p; g.lda = np;
        g.B = m->U.p; g.ldb = np;
        g.M = g.N = (int)np; g.K = (int)np;
        g.alpha = 1.0; g.beta_one = 0; g.lower = 1; g.k_from_tile = 1;
        bool done = false;
        if (tc) {
            // on tcgen05: the rows of U in 8 digit slices (np x np x 8 bytes), exact int8 products accumulated in int32 over the whole
            // contraction (8 np 65 64 < 2^31 up to np = 32768), added to a zeroed K^-1
            if (m->ozU.reserve((size_t)np * np) == cudaSuccess && m->ozScale.reserve((size_t)std::max<int64_t>(np, m->cap)) == cudaSuccess) {
                int8_t* dg = reinterpret_cast<int8_t*>(m->ozU.p);
                ozaki_slice_launch(m->U.p, np, np, (int)np, dg, m->ozScale.p, m->ctx());
                CU(m, cudaMemsetAsync(m->Kinv.p, 0, (size_t)np * np * sizeof(double), m->st));
                g.beta_one = 1;
                m->launches += 2 + (ozaki_update_launch(g, dg, m->ozScale.p, dg, m->ozScale.p, 0, m->ctx()) > 0);
                done = true;
            } else {
                cudaGetLastError();
            }
        }
        if (!done) m->launches += gemm_nt_launch(g, m->ctx()) > 0;
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

// ---------------------------------------------------------------------------------------------------------------------
// The same gradient with the O(n^3) work spread over the ranks of the communicator (every rank holds the full factor after
// fgp_fit_sharded):
//   1. U = L^-T by ROWS: the rows of the forward solve on the identity are independent, rank r takes the block rows
//      r, r+P, ... (cyclic: equal shares of the triangle), n^3 / (3 P) flops and no communication;
//   2. ncclAllGather of the row shares (8 n^2 / 2 bytes in total) + a local un-permutation into the natural layout;
//   3. K^-1 = U U^T on the tile columns of the 512-column panels the rank owns in the fit's distribution (one lower-mode
//      launch, k_from_tile), n^3 / (3 P) flops;
//   4. the pair-tile reductions over the owned columns, then ONE ncclAllReduce of the 2 P + 1 partial sums.
__global__ void set_cyclic_identity_kernel(double* Xt, int64_t ld, int64_t nblocks, int64_t first, int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local row
    if (i >= nblocks * TILE) return;
    const int64_t t = i / TILE, b = first + t * stride;
    Xt[i + (b * TILE + i % TILE) * ld] = 1.0;
}
// U[(b*128 + i), c] = Ug[rank b % P][(b / P)*128 + i, c]  (chunk leading dimension mmax)
__global__ void unpermute_rows_kernel(const double* __restrict__ Ug, int64_t mmax, int64_t np, int P, double* __restrict__ U) {
    const int64_t c = blockIdx.y;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < np; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = row / TILE, i = row % TILE;
        U[row + c * np] = Ug[(size_t)(b % P) * mmax * np + (b / P) * TILE + i + c * mmax];
    }
}

int lml_gradient_sharded_device(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int scaled,
                                double* scale_out, double* grads) {
    fgp_comm* cm = m->comm;
    const int P = cm->nranks, r = cm->rank;
    // (a communicator of ONE rank runs the very same steps with the two collectives replaced by a copy / nothing, so that the
    // schedule is testable on a single GPU)
    const NcclApi* nccl = (P > 1) ? nccl_api() : nullptr;
    if (P > 1 && !nccl) return fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
    const int64_t np = m->np, nb = np / TILE;
    const int np_params = kt.nparams;
    const int64_t tmax = (nb + P - 1) / P, mmax = tmax * TILE;            // block rows per rank (padded to the largest share)
    const int64_t mine = (nb > r) ? (nb - r + P - 1) / P : 0;              // block rows r, r+P, ... < nb
    // 1. my rows of U (mmax x np, ld mmax); rows beyond my share stay zero
    CU(m, m->lml_rows.reserve((size_t)mmax * np));
    CU(m, m->U.reserve((size_t)np * np));
    CU(m, m->Kinv.reserve((size_t)np * np + (size_t)P * mmax * np));      // K^-1, then the gather buffer
    double* Ug = m->Kinv.p + (size_t)np * np;
    CU(m, cudaMemsetAsync(m->lml_rows.p, 0, (size_t)mmax * np * sizeof(double), m->st));
    if (mine > 0)
        set_cyclic_identity_kernel<<<(unsigned)((mine * TILE + 255) / 256), 256, 0, m->st>>>(m->lml_rows.p, mmax, mine, r, P);
    m->launches += 1;
    m->launches += trsm_fwd_t(m->lml_rows.p, mmax, mine * TILE, m->L.p, m->cap, m->inv.p, 0, nb, nullptr, m->ctx(), true, r, P);
    // 2. all ranks' rows, then the natural layout
    if (P > 1) {
        if (nccl->AllGather(m->lml_rows.p, Ug, (size_t)mmax * np, ncclDouble, cm->comm, m->st) != ncclSuccess)
            return fail(m, FGP_ERR_COMM, "ncclAllGather of the rows of L^-T failed");
    } else {
        CU(m, cudaMemcpyAsync(Ug, m->lml_rows.p, (size_t)mmax * np * sizeof(double), cudaMemcpyDeviceToDevice, m->st));
    }
    unpermute_rows_kernel<<<dim3(8, (unsigned)np), 256, 0, m->st>>>(Ug, mmax, np, P, m->U.p);
    m->launches += 1;
    // 3. K^-1 on the owned panels: tile columns [4 p, 4 p + 4) for p = r, r+P, ...
    const int64_t PT = HEAD_PANEL / TILE, NPn = (nb + PT - 1) / PT;
    int64_t ncols = 0;
    for (int64_t p = r; p < NPn; p += P) ncols += std::min<int64_t>(PT, nb - p * PT);
    if (ncols > 0) {
        const int64_t c0 = (int64_t)r * PT;
        GemmArgs g{};
        g.C = m->Kinv.p + c0 * TILE * (np + 1); g.ldc = np;
        g.A = m->U.p + c0 * TILE; g.lda = np;
        g.B = g.A; g.ldb = np;
        g.M = (int)(np - c0 * TILE); g.N = (int)(ncols * TILE); g.K = (int)np;
        g.alpha = 1.0; g.beta_one = 0; g.lower = 1; g.k_from_tile = 1; g.k_tile0 = (int)c0;
        g.grp = (int)PT; g.stride = (int)(P * PT);
        m->launches += gemm_nt_launch(g, m->ctx()) > 0;
    }
    // 4. reductions over the owned columns: one pair-tile launch per owned panel, partials laid out one after the other
    const DevKernel dk = to_dev(kd);
    const int mode = (kt.need_d2 ? PAIR_D2 : 0) | (kt.need_dot ? PAIR_DOT : 0);
    const bool fast = (kt.kind == KIND_SQEXP || kt.kind == KIND_MATERN2);
    const int pmax = fast ? 2 : FGP_MAX_PARAMS, nv = 2 * pmax + 1;
    int64_t blocks = 0;
    for (int64_t p = r; p < NPn; p += P) {
        const int64_t w = std::min<int64_t>(PT, nb - p * PT);
        blocks += (nb - p * PT) * (w * TILE / PAIR_TN);
    }
    CU(m, m->lml_partial.reserve((size_t)std::max<int64_t>(blocks, 1) * nv));
    int64_t off = 0;
    for (int64_t p = r; p < NPn; p += P) {
        const int64_t w = std::min<int64_t>(PT, nb - p * PT);
        PairArgs pa{};
        pa.xa_c = pa.xb_c = m->xc.p;
        pa.xa_r = pa.xb_r = m->xr.p;
        pa.na = pa.nb = m->nc.p;
        pa.dp = (int)m->dp;
        pa.rows = pa.cols = np;
        pa.row_tile0 = (int)(p * PT);
        pa.col_tile0 = (int)(p * PT * TILE / PAIR_TN);
        pa.col_tiles = (int)(w * TILE / PAIR_TN);
        pa.symmetric = 1;
        double* part = m->lml_partial.p + off * nv;
        if (kt.kind == KIND_SQEXP) launch_lml<KIND_SQEXP, 2, PAIR_D2>(m, pa, dk, m->Kinv.p, np, part);
        else if (kt.kind == KIND_MATERN2) launch_lml<KIND_MATERN2, 2, PAIR_D2>(m, pa, dk, m->Kinv.p, np, part);
        else if (mode == PAIR_D2) launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_D2>(m, pa, dk, m->Kinv.p, np, part);
        else if (mode == PAIR_DOT) launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_DOT>(m, pa, dk, m->Kinv.p, np, part);
        else launch_lml<KIND_GENERIC, FGP_MAX_PARAMS, PAIR_BOTH>(m, pa, dk, m->Kinv.p, np, part);
        const dim3 grid = pair_grid(pa);
        off += (int64_t)grid.x * grid.y;
    }
    CU(m, m->scalars.reserve(64));
    CU(m, cudaMemsetAsync(m->scalars.p, 0, 64 * sizeof(double), m->st));
    if (off > 0) partials_final_kernel<<<1, 64, 0, m->st>>>(m->lml_partial.p, off, nv, m->scalars.p);
    if (P > 1 && nccl->AllReduce(m->scalars.p, m->scalars.p, (size_t)nv, ncclDouble, ncclSum, cm->comm, m->st) != ncclSuccess)
        return fail(m, FGP_ERR_COMM, "ncclAllReduce of the gradient sums failed");
    reduce_kernel<2><<<1, 256, 0, m->st>>>(m->alpha.p, m->alpha.p, m->n, dk, 0.0, m->scalars.p + 50);  // alpha.alpha
    reduce_kernel<2><<<1, 256, 0, m->st>>>(m->y.p, m->alpha.p, m->n, dk, 0.0, m->scalars.p + 51);      // y.alpha
    m->launches += 3;
    FGP_TRY(ensure_pinned(m, 64));
    CU(m, cudaMemcpyAsync(m->pinned, m->scalars.p, 52 * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    const double* h = m->pinned;
    const double scale = scaled ? h[51] / (double)m->n : 1.0;  // optimizer.rs:174
    for (int p = 0; p < np_params; ++p) {
        double data_fit = h[pmax + p];
        if (scaled) data_fit /= scale;                          // optimizer.rs:186
        grads[p] = (data_fit - h[p]) / 2.0;                     // optimizer.rs:192 / :49
    }
    if (!scaled) grads[np_params] = noise * (h[50] - h[2 * pmax]);  // optimizer.rs:54-57
    if (scale_out) *scale_out = scale;
    m->kinv_valid = false;  // only this rank's panels of K^-1 are on the device
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

// ---------------------------------------------------------------------------------------------------------------------
// LinearPrior::fit (src/parameters/prior.rs:139-159: least squares of y on [1 | X]).  On the CENTRED resident points Xc the
// intercept decouples (the columns of Xc sum to zero): slopes w solve (Xc^T Xc) w = Xc^T y, intercept = mean(y) - mean(X).w.
// One kernel accumulates, per block of rows, the d x (d + 2) table [Xc^T Xc | Xc^T y | .] and sum(y); partials are summed in a
// fixed order (deterministic); the d x d system is solved on the host (the reference's SVD solve is host code too).
__global__ void __launch_bounds__(256) prior_normal_kernel(const double* __restrict__ xc, int dp, int d, const double* __restrict__ y,
                                                           int64_t n, int64_t rows_per_block, double* __restrict__ partial) {
    extern __shared__ double sm[];   // [256][d + 1]: centred point, y
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
    const int W = d + 1, nout = d * W + 1;
    double acc[8];   // this thread's outputs t = threadIdx.x + 256 k (nout <= 8 * 256)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0;
    for (int64_t c0 = r0; c0 < r1; c0 += 256) {
        const int64_t r = c0 + threadIdx.x;
        if (r < r1) {
            for (int j = 0; j < d; ++j) sm[threadIdx.x * W + j] = xc[r * dp + j];
            sm[threadIdx.x * W + d] = y[r];
        } else {
            for (int j = 0; j <= d; ++j) sm[threadIdx.x * W + j] = 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = threadIdx.x + 256 * k;
            if (t < nout) {
                double a = acc[k];
                if (t == nout - 1) {
                    for (int q = 0; q < 256; ++q) a += sm[q * W + d];                       // sum(y)
                } else {
                    const int i = t / W, j = t % W;                                          // j == d: Xc^T y
                    for (int q = 0; q < 256; ++q) a = fma(sm[q * W + i], sm[q * W + j], a);
                }
                acc[k] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int t = threadIdx.x + 256 * k;
        if (t < nout) partial[(int64_t)blockIdx.x * nout + t] = acc[k];
    }
}
__global__ void prior_normal_final_kernel(const double* __restrict__ partial, int blocks, int nout, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout) return;
    double a = 0.0;
    for (int b = 0; b < blocks; ++b) a += partial[(int64_t)b * nout + t];
    out[t] = a;
}

int linear_prior_fit_device(fgp_model* m, const double* y_host, double* weights, double* intercept) {
    const int d = (int)m->d, W = d + 1, nout = d * W + 1;
    if (nout > 8 * 256) return FGP_ERR_BAD_ARG;   // d <= 44
    const int64_t n = m->n, rpb = 2048;
    const int blocks = (int)((n + rpb - 1) / rpb);
    CU(m, m->work.reserve((size_t)std::max<int64_t>(m->cap, n)));
    CU(m, cudaMemcpyAsync(m->work.p, y_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, m->st));
    CU(m, m->lml_partial.reserve((size_t)blocks * nout + nout + d));
    double* out = m->lml_partial.p + (size_t)blocks * nout;
    const size_t smem = (size_t)256 * W * sizeof(double);
    if (smem > 48 * 1024) CU(m, cudaFuncSetAttribute(prior_normal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prior_normal_kernel<<<blocks, 256, smem, m->st>>>(m->xc.p, (int)m->dp, d, m->work.p, n, rpb, m->lml_partial.p);
    prior_normal_final_kernel<<<(nout + 127) / 128, 128, 0, m->st>>>(m->lml_partial.p, blocks, nout, out);
    m->launches += 2;
    CU(m, cudaMemcpyAsync(out + nout, m->cmean.p, (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice, m->st));
    FGP_TRY(ensure_pinned(m, (size_t)nout + d));
    CU(m, cudaMemcpyAsync(m->pinned, out, (size_t)(nout + d) * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    // host: Cholesky of the d x d normal matrix G (row i of the table: G[i][0..d), then (Xc^T y)[i])
    std::vector<double> G((size_t)d * d), b((size_t)d);
    for (int i = 0; i < d; ++i) {
        for (int j = 0; j < d; ++j) G[(size_t)i * d + j] = m->pinned[i * W + j];
        b[i] = m->pinned[i * W + d];
    }
    for (int j = 0; j < d; ++j) {
        double s = G[(size_t)j * d + j];
        for (int k = 0; k < j; ++k) s -= G[(size_t)j * d + k] * G[(size_t)j * d + k];
        if (!(s > 0.0)) return fail(m, FGP_ERR_NOT_POSDEF, "linear prior fit: the centred inputs are rank deficient");
        const double ljj = std::sqrt(s);
        G[(size_t)j * d + j] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double t = G[(size_t)i * d + j];
            for (int k = 0; k < j; ++k) t -= G[(size_t)i * d + k] * G[(size_t)j * d + k];
            G[(size_t)i * d + j] = t / ljj;
        }
    }
    for (int i = 0; i < d; ++i) {   // L z = b
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= G[(size_t)i * d + k] * b[k];
        b[i] = t / G[(size_t)i * d + i];
    }
    for (int i = d - 1; i >= 0; --i) {   // L^T w = z
        double t = b[i];
        for (int k = i + 1; k < d; ++k) t -= G[(size_t)k * d + i] * b[k];
        b[i] = t / G[(size_t)i * d + i];
    }
    double c = m->pinned[nout - 1] / (double)n;   // mean(y)
    for (int i = 0; i < d; ++i) {
        weights[i] = b[i];
        c -= m->pinned[nout + i] * b[i];          // - mean(X) . w
    }
    *intercept = c;
    return FGP_OK;
}

void launch_set_identity(double* A, int64_t ld, int64_t n, cudaStream_t st) {
    set_identity_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A, ld, n);
}

}  // namespace fgp
