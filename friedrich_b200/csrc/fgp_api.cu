// fgp_api.cu — C-ABI entry points of libfgp_sm100.so (see include/fgp.h for the contract and the reference
// call sites each function replaces).  One handle = one GPU, one stream; there is no CPU fallback anywhere:
// every numeric result below is produced by the CUDA kernels in this directory or the call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/fgp.h"
#include "common.cuh"
#include "gemm_nt.cuh"
#include "kernel_eval.cuh"
#include "model.cuh"
#include "pair_tiles.cuh"
#include "potrf.cuh"
#include "vector_kernels.cuh"
#include "trsm.cuh"
#include "lml.cuh"
#include "covariance.cuh"
#include "sharded.cuh"
#include "ozaki.cuh"

using namespace fgp;

#define FGP_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

int comm_agree(fgp_model* m, int local_rc);  // cross-rank status exchange before a collective part (defined with the sharded entry points)

// upload a column-major host matrix (rows x cols, ld) compactly into m->staging (ld = rows)
int upload_colmajor(fgp_model* m, const double* src, int64_t ld, int64_t rows, int64_t cols) {
    CU(m, m->staging.reserve((size_t)rows * cols));
    CU(m, cudaMemcpy2DAsync(m->staging.p, rows * sizeof(double), src, ld * sizeof(double), rows * sizeof(double), cols,
                            cudaMemcpyHostToDevice, m->st));
    return FGP_OK;
}

PairArgs train_pair_args(const fgp_model* m) {
    PairArgs pa{};
    pa.xa_c = pa.xb_c = m->xc.p;
    pa.xa_r = pa.xb_r = m->xr.p;
    pa.na = pa.nb = m->nc.p;
    pa.dp = (int)m->dp;
    pa.rows = pa.cols = m->np;
    pa.row_tile0 = 0;
    pa.symmetric = 1;
    return pa;
}

// ---- alpha = K^-1 y : z = L^-1 y (kept for the likelihood, mod.rs:203), alpha = L^-T z --------------------------
void solve_alpha(fgp_model* m) {
    // z = L^-1 y, alpha = L^-T z: one wavefront launch each (csrc/vector_kernels.cuh); flags[0..nb) forward, [nb..2nb) adjoint
    const int nb = (int)(m->np / TILE);
    int* flags = reinterpret_cast<int*>(m->work.p);  // np doubles of scratch >= 2 nb flags + 2 tickets
    cudaMemsetAsync(flags, 0, (2 * (size_t)nb + 2) * sizeof(int), m->st);
    static bool attr_done_dev[64] = {};
    bool& attr_done = *per_device_flag(attr_done_dev);
    if (!attr_done) {  // the block's inverse diagonal tile lives in 128 KiB of dynamic shared memory
        cudaFuncSetAttribute(trsv_fwd_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_WAVE_SMEM);
        cudaFuncSetAttribute(trsv_adj_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_WAVE_SMEM);
        attr_done = true;
    }
    trsv_fwd_wave_kernel<<<nb, TRSV_THREADS, TRSV_WAVE_SMEM, m->st>>>(m->L.p, m->cap, m->inv.p, m->y.p, m->z.p, flags,
                                                                      flags + 2 * nb);
    trsv_adj_wave_kernel<<<nb, TRSV_THREADS, TRSV_WAVE_SMEM, m->st>>>(m->L.p, m->cap, m->invT.p, m->z.p, m->alpha.p, flags + nb,
                                                                      flags + 2 * nb + 1, nb);
    m->launches += 2;
}

// The per-panel inverse blocks W are only read inside the factorisation that produces them, so a restored factor
// (fgp_upload_state) needs none: the panel table is simply emptied.
int rebuild_panel_inverses(fgp_model* m) {
    m->pstart.clear();
    m->w_valid = false;
    m->ozL_valid = false;
    return FGP_OK;
}

// Buffers of the head schedule for factoring block columns [jb_begin, np/128): the panel table keeps the panels that start
// before jb_begin (a panel cut by jb_begin stays valid: the leading block of a triangular inverse is the inverse of the
// leading block), the new ones follow.  *p0 = W / sync slot of the first new panel.
int prepare_head_work(fgp_model* m, int64_t jb_begin, PotrfWork* w, int64_t* p0) {
    const int64_t nb = m->np / TILE, PT = HEAD_PANEL / TILE;
    while (!m->pstart.empty() && m->pstart.back() >= jb_begin) m->pstart.pop_back();
    *p0 = (int64_t)m->pstart.size();
    for (int64_t J = jb_begin; J < nb; J += PT) m->pstart.push_back(J);
    const int64_t slots = (int64_t)m->pstart.size();
    CU(m, m->Wp.reserve((size_t)slots * HEAD_PANEL * HEAD_PANEL, true, m->st, true));  // diagonal tiles: lower triangles written only
    CU(m, m->Pscr.reserve((size_t)HEAD_PANEL * HEAD_PANEL));
    CU(m, m->pbuf[0].reserve((size_t)m->cap * HEAD_PANEL));
    CU(m, m->pbuf[1].reserve((size_t)m->cap * HEAD_PANEL));
    if (slots > m->head_sync_cap) {
        if (m->head_sync) cudaFree(m->head_sync);
        m->head_sync = nullptr;
        m->head_sync_cap = 0;
        const int64_t cap = std::max<int64_t>(2 * slots, 64);
        CU(m, cudaMalloc(&m->head_sync, (size_t)cap * HEAD_SYNC_INTS * sizeof(int)));
        m->head_sync_cap = cap;
    }
    *w = PotrfWork{m->inv.p, m->invT.p, m->Wp.p, m->Pscr.p, m->head_sync, {m->pbuf[0].p, m->pbuf[1].p}, m->evTop, m->evRest,
                   {m->evCopy[0], m->evCopy[1]}};
    if (m->tcgen05) {   // 8 digit bytes per panel element = as many bytes as the f64 panel buffer
        CU(m, m->ozDigits.reserve((size_t)m->cap * HEAD_PANEL));
        CU(m, m->ozScale.reserve((size_t)m->cap));
        if (ozaki_prepare() != cudaSuccess) return fail(m, FGP_ERR_CUDA, "tcgen05 update kernel could not be configured");
        w->oz_digits = reinterpret_cast<int8_t*>(m->ozDigits.p);
        w->oz_scale = m->ozScale.p;
        m->ozL_valid = false;
        if (jb_begin == 0) {
            // full fit (single GPU, or sharded: every rank slices every received panel): every panel with >= OZ_MIN_ROWS rows below it keeps its digit slices (8 bytes per element of
            // L below the panel's diagonal block: as many bytes as that part of L) for the solves of predict
            m->ozOffBytes.assign((size_t)slots, -1);
            m->ozOffRows.assign((size_t)slots, -1);
            int64_t bytes = 0, rows_total = 0;
            for (int64_t s = 0; s < slots; ++s) {
                const int64_t J = m->pstart[s], Jend = std::min<int64_t>(J + PT, nb), rows = m->np - Jend * TILE;
                if (rows < OZ_MIN_ROWS) continue;
                m->ozOffBytes[s] = bytes;
                m->ozOffRows[s] = rows_total;
                bytes += (int64_t)ozaki_slice_bytes(rows, (int)((Jend - J) * TILE));
                rows_total += rows;
            }
            if (bytes > 0) {
                CU(m, m->ozL.reserve((size_t)bytes / 8));
                CU(m, m->ozLscale.reserve((size_t)rows_total));
                w->oz_digits = reinterpret_cast<int8_t*>(m->ozL.p);
                w->oz_scale = m->ozLscale.p;
                w->oz_off_bytes = m->ozOffBytes.data();
                w->oz_off_rows = m->ozOffRows.data();
            }
        }
    } else {
        m->ozL_valid = false;
    }
    return FGP_OK;
}

int reserve_training(fgp_model* m, int64_t cap_rows, int64_t dp, bool keep) {
    // `keep` is used by add_samples when the capacity grows: point arrays and vectors keep their prefix; L is handled by
    // the caller (its leading dimension changes).
    CU(m, m->xr.reserve((size_t)cap_rows * dp, keep, m->st));
    CU(m, m->xc.reserve((size_t)cap_rows * dp, keep, m->st));
    CU(m, m->nc.reserve((size_t)cap_rows, keep, m->st));
    CU(m, m->nr.reserve((size_t)cap_rows, keep, m->st));
    CU(m, m->cmean.reserve((size_t)std::max<int64_t>(dp, 8), keep, m->st));
    CU(m, m->y.reserve((size_t)cap_rows, keep, m->st));
    CU(m, m->z.reserve((size_t)cap_rows));
    CU(m, m->alpha.reserve((size_t)cap_rows));
    CU(m, m->work.reserve((size_t)cap_rows));
    CU(m, m->inv.reserve((size_t)cap_rows * TILE, keep, m->st, true));  // the head kernel writes the lower triangles only
    CU(m, m->invT.reserve((size_t)cap_rows * TILE, keep, m->st));
    return FGP_OK;
}

// Gram lower triangle + noise^2 I (algebra/mod.rs:67-79), blocked Cholesky (algebra/mod.rs:81-91), alpha.
int factor_resident(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps,
                    double eps) {
    CU(m, cudaMemsetAsync(m->info_d, 0, sizeof(int), m->st));
    // Gram: the first panel's block columns first; the columns behind it are assembled on the main stream WHILE the panel
    // stream already factors the first panel (potrf_lower's hook)
    const int64_t nb = m->np / TILE, pt = std::min<int64_t>(panel_tiles(m->np), nb);
    PairArgs pa0 = train_pair_args(m);
    pa0.col_tile0 = 0;
    pa0.col_tiles = (int)(pt * TILE / PAIR_TN);
    write_covariance(m, kt, kd, pa0, m->L.p, m->cap, m->n, m->n, noise * noise);
    const std::function<void()> rest = [&]() {
        if (pt >= nb) return;
        PairArgs pa1 = train_pair_args(m);
        pa1.row_tile0 = (int)pt;  // rows above the first remaining diagonal tile are in the upper triangle
        pa1.col_tile0 = (int)(pt * TILE / PAIR_TN);
        pa1.col_tiles = (int)((nb - pt) * TILE / PAIR_TN);
        write_covariance(m, kt, kd, pa1, m->L.p, m->cap, m->n, m->n, noise * noise);
    };
    PotrfCounters cnt;
    const PotrfLookahead la{m->st2, m->evA, m->evB, m->st3, m->evC, m->evD};
    if (m->head_schedule) {
        PotrfWork w;
        int64_t p0 = 0;
        FGP_TRY(prepare_head_work(m, 0, &w, &p0));
        potrf_lower_head(m->L.p, m->cap, m->np, 0, w, p0, has_eps, eps, m->info_d, m->ctx(), m->lookahead ? &la : nullptr, &cnt,
                         &rest);
        m->w_valid = true;
        m->ozL_valid = w.oz_off_bytes != nullptr;
    } else {
        potrf_lower(m->L.p, m->cap, m->np, 0, m->inv.p, m->invT.p, has_eps, eps, m->info_d, m->ctx(),
                    m->lookahead ? &la : nullptr, &cnt, &rest);
        m->pstart.clear();
        m->w_valid = false;
    }
    m->launches += cnt.launches;
    CU(m, cudaMemcpyAsync(m->info_h, m->info_d, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    solve_alpha(m);
    CU(m, cudaStreamSynchronize(m->st));
    CU(m, cudaGetLastError());
    if (*m->info_h == HEAD_TIMEOUT) {
        m->fitted = false;
        return fail(m, FGP_ERR_CUDA, "internal error: a dependency wait inside potrf_head_kernel timed out");
    }
    if (*m->info_h != 0) {
        m->failed_col = *m->info_h - 1;
        m->fitted = false;
        return fail(m, FGP_ERR_NOT_POSDEF,
                    "Cholesky decomposition failed at column " + std::to_string(m->failed_col));
    }
    m->failed_col = -1;
    m->fitted = true;
    m->kinv_valid = false;
    return FGP_OK;
}

int check_kernel(fgp_model* m, const fgp_kernel_desc* kd, KernelTraits* kt) {
    *kt = classify(kd);
    if (!kt->valid) return fail(m, FGP_ERR_BAD_KERNEL, "malformed or unsupported kernel descriptor");
    return FGP_OK;
}

int stage_queries(fgp_model* m, const double* Xq, int64_t ldq, int64_t q) {
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!Xq || q <= 0 || ldq < q) return fail(m, FGP_ERR_BAD_ARG, "bad query matrix");
    const int64_t qp = round_up(q, TILE);
    FGP_TRY(upload_colmajor(m, Xq, ldq, q, m->d));
    CU(m, m->qr.reserve((size_t)qp * m->dp));
    CU(m, m->qc.reserve((size_t)qp * m->dp));
    CU(m, m->qnc.reserve((size_t)qp));
    CU(m, m->qnr.reserve((size_t)qp));
    CU(m, m->bt.reserve((size_t)qp * m->np));
    CU(m, m->partial.reserve((size_t)std::max<int64_t>(m->np / ROWRED_CHUNK, 2) * qp));
    CU(m, m->mean_d.reserve((size_t)qp));
    CU(m, m->var_d.reserve((size_t)qp));
    convert_points_kernel<<<(unsigned)((qp + 255) / 256), 256, 0, m->st>>>(m->staging.p, q, q, (int)m->d, (int)m->dp,
                                                                          m->cmean.p, 0, qp, m->qr.p, m->qc.p, m->qnc.p,
                                                                          m->qnr.p);
    m->launches += 1;
    m->q = q;
    m->qp = qp;
    m->have_mean = m->have_var = false;
    return FGP_OK;
}

PairArgs query_pair_args(const fgp_model* m) {
    PairArgs pa{};
    pa.xa_c = m->qc.p;
    pa.xa_r = m->qr.p;
    pa.na = m->qnc.p;
    pa.xb_c = m->xc.p;
    pa.xb_r = m->xr.p;
    pa.nb = m->nc.p;
    pa.dp = (int)m->dp;
    pa.rows = m->qp;
    pa.cols = m->np;
    pa.row_tile0 = 0;
    pa.symmetric = 0;
    return pa;
}

// Device part of predict / predict_variance / predict_mean_variance on the staged queries.
//   Bt (qp x np) = k(query, train)                     make_covariance_matrix, algebra/mod.rs:41-54 (transposed)
//   mean = Bt * alpha                                   == (K^-1 K_nq)^T y  (mod.rs:235,241) with alpha = K^-1 y cached
//   Bt <- Bt * L^-T                                     l().solve_lower_triangular (mod.rs:260-263), transposed
//   var_i = k(q_i,q_i) - || Bt[i,:] ||^2                mod.rs:266-270
// Latency path for a handful of queries (the Bayesian-optimisation use case: one candidate at a time): the 128-step chain of
// the multi-RHS solve costs ~10 ms however few right-hand sides there are, a wavefront forward substitution per query 0.4 ms.
//   Kc (np x q, one query per COLUMN) = k(train, query);  mean_i = Kc[:, i] . alpha;  z_i = L^-1 Kc[:, i] in place;
//   var_i = k(q_i, q_i) - ||z_i||^2                                                            (mod.rs:235-241, :260-270)
constexpr int64_t PREDICT_SMALL_Q = 16;  // up to here ONE multi-right-hand-side wavefront launch; above, the tensor-pipe path (2.5 ms at n = 16384)
static_assert(PREDICT_SMALL_Q <= TRSV_MULTI_QMAX, "latency path of predict: right-hand sides per wavefront launch");
int predict_small(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, int want_mean, int want_var) {
    const int64_t qp = m->qp, np = m->np, q = m->q;
    const int nb = (int)(np / TILE);
    PairArgs pa{};
    pa.xa_c = m->xc.p; pa.xa_r = m->xr.p; pa.na = m->nc.p;     // rows: training points
    pa.xb_c = m->qc.p; pa.xb_r = m->qr.p; pa.nb = m->qnc.p;    // columns: queries
    pa.dp = (int)m->dp;
    pa.rows = np;
    pa.cols = qp;
    pa.symmetric = 0;
    double* Kc = m->bt.p;  // np x qp, ld = np (the buffer holds qp x np doubles)
    write_covariance(m, kt, kd, pa, Kc, np, m->n, q, 0.0);
    const DevKernel dk = to_dev(kd);
    if (want_mean) {
        col_reduce_kernel<0><<<(unsigned)q, 256, 0, m->st>>>(Kc, np, m->alpha.p, np, m->partial.p);
        rowreduce_final_kernel<<<(unsigned)(qp / 128), 128, 0, m->st>>>(m->partial.p, 1, qp, q, 0, dk, m->qnr.p, m->mean_d.p);
        m->launches += 2;
        m->have_mean = true;
    }
    if (want_var) {
        int* flags = reinterpret_cast<int*>(m->work.p);  // np doubles of scratch >= nb flags + 1 ticket
        CU(m, cudaMemsetAsync(flags, 0, (size_t)(nb + 1) * sizeof(int), m->st));
        static bool attr_done_dev[64] = {};
        bool& attr_done = *per_device_flag(attr_done_dev);
        if (!attr_done) {
            cudaFuncSetAttribute(trsv_fwd_wave_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_MULTI_SMEM);
            attr_done = true;
        }
        // ONE wavefront launch for all the right-hand sides (a launch per query costs q x 0.45 ms at n = 16384)
        if (q == 1)
            trsv_fwd_wave_kernel<<<nb, TRSV_THREADS, TRSV_WAVE_SMEM, m->st>>>(m->L.p, m->cap, m->inv.p, Kc, Kc, flags, flags + nb);
        else
            trsv_fwd_wave_multi_kernel<<<nb, TRSV_THREADS, TRSV_MULTI_SMEM, m->st>>>(m->L.p, m->cap, m->inv.p, Kc, np, (int)q,
                                                                                   flags, flags + nb);
        col_reduce_kernel<1><<<(unsigned)q, 256, 0, m->st>>>(Kc, np, nullptr, np, m->partial.p + qp);
        rowreduce_final_kernel<<<(unsigned)(qp / 128), 128, 0, m->st>>>(m->partial.p + qp, 1, qp, q, 1, dk, m->qnr.p, m->var_d.p);
        m->launches += 3;
        m->have_var = true;
    }
    return FGP_OK;
}

int predict_device(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, int want_mean, int want_var) {
    const int64_t qp = m->qp, np = m->np;
    if (want_var && m->q <= PREDICT_SMALL_Q && !m->force_batched_predict) return predict_small(m, kd, kt, want_mean, want_var);
    write_covariance(m, kt, kd, query_pair_args(m), m->bt.p, qp, m->q, m->n, 0.0);
    const dim3 rgrid((unsigned)(qp / 128), (unsigned)(np / ROWRED_CHUNK));
    const int chunks = (int)(np / ROWRED_CHUNK);
    const DevKernel dk = to_dev(kd);
    if (want_mean) {
        rowreduce_partial_kernel<0><<<rgrid, 128, 0, m->st>>>(m->bt.p, qp, m->alpha.p, m->partial.p, qp);
        rowreduce_final_kernel<<<(unsigned)(qp / 128), 128, 0, m->st>>>(m->partial.p, chunks, qp, m->q, 0, dk, m->qnr.p,
                                                                         m->mean_d.p);
        m->launches += 2;
        m->have_mean = true;
    }
    if (want_var) {
        if (m->w_valid && !m->pstart.empty() && qp <= m->cap)  // panels of the head schedule: two K <= 512 launches per 512 columns
        {
            // updates behind a panel whose digit slices the fit kept run on tcgen05 (the solved panel of Bt is sliced on the fly)
            OzPanelStore ozs{};
            const bool use_oz = m->tcgen05 && m->ozL_valid && m->ozOffBytes.size() == m->pstart.size() && qp <= m->cap && m->ozDigits.p;
            if (use_oz)
                ozs = OzPanelStore{reinterpret_cast<const int8_t*>(m->ozL.p), m->ozLscale.p, m->ozOffBytes.data(), m->ozOffRows.data(),
                                   reinterpret_cast<int8_t*>(m->ozDigits.p), m->ozScale.p};
            m->launches += trsm_fwd_t_panels(m->bt.p, qp, qp, m->L.p, m->cap, m->Wp.p, m->pstart.data(), (int64_t)m->pstart.size(),
                                             np / TILE, np / TILE, m->pbuf[0].p, m->pbuf[1].p, m->ctx(),
                                             m->lookahead ? m->st2 : nullptr, m->evA, m->evB, m->evC, use_oz ? &ozs : nullptr);
        }
        else if (m->lookahead)
            m->launches += trsm_fwd_t_lookahead(m->bt.p, qp, qp, m->L.p, m->cap, m->inv.p, np / TILE, m->ctx(), m->st2, m->evA,
                                                m->evB);
        else
            m->launches += trsm_fwd_t(m->bt.p, qp, qp, m->L.p, m->cap, m->inv.p, 0, np / TILE, nullptr, m->ctx());
        rowreduce_partial_kernel<1><<<rgrid, 128, 0, m->st>>>(m->bt.p, qp, nullptr, m->partial.p, qp);
        rowreduce_final_kernel<<<(unsigned)(qp / 128), 128, 0, m->st>>>(m->partial.p, chunks, qp, m->q, 1, dk, m->qnr.p,
                                                                         m->var_d.p);
        m->launches += 2;
        m->have_var = true;
    }
    return FGP_OK;
}

int fetch_predictions(fgp_model* m, double* mean, double* var) {
    FGP_TRY(ensure_pinned(m, (size_t)2 * m->qp));
    if (mean) {
        if (!m->have_mean) return fail(m, FGP_ERR_BAD_ARG, "mean was not computed");
        CU(m, cudaMemcpyAsync(m->pinned, m->mean_d.p, m->q * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    }
    if (var) {
        if (!m->have_var) return fail(m, FGP_ERR_BAD_ARG, "variance was not computed");
        CU(m, cudaMemcpyAsync(m->pinned + m->qp, m->var_d.p, m->q * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    }
    CU(m, cudaStreamSynchronize(m->st));
    if (mean) std::memcpy(mean, m->pinned, m->q * sizeof(double));
    if (var) std::memcpy(var, m->pinned + m->qp, m->q * sizeof(double));
    return FGP_OK;
}

int predict_common(fgp_model* m, const fgp_kernel_desc* kd, const double* Xq, int64_t ldq, int64_t q, double* mean,
                   double* var) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kd, &kt));
    if (!mean && !var) return fail(m, FGP_ERR_BAD_ARG, "no output buffer");
    begin_timed(m);
    FGP_TRY(stage_queries(m, Xq, ldq, q));
    FGP_TRY(predict_device(m, kd, kt, mean != nullptr, var != nullptr));
    FGP_TRY(end_timed(m));
    return fetch_predictions(m, mean, var);
}

}  // namespace

namespace {
void comm_release(fgp_model* m) {
    fgp_comm* c = m->comm;
    if (!c) return;
    if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
    c->pbuf[0].release();
    c->pbuf[1].release();
    if (c->st_comm) {
        cudaStreamSynchronize(c->st_comm);
        cudaStreamDestroy(c->st_comm);
    }
    if (c->st_copy) {
        cudaStreamSynchronize(c->st_copy);
        cudaStreamDestroy(c->st_copy);
    }
    if (c->ev_col) cudaEventDestroy(c->ev_col);
    if (c->ev_bcast) cudaEventDestroy(c->ev_bcast);
    for (cudaEvent_t e : c->ev_trail)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_copy)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pipe)
        if (e) cudaEventDestroy(e);
    delete c;
    m->comm = nullptr;
}
}  // namespace

// =================================================================================================================
// lifecycle
FGP_EXPORT const char* fgp_version(void) { return "libfgp_sm100 0.1 (sm_100a, fp64 DMMA + TMA)"; }

FGP_EXPORT int fgp_create(int device, fgp_model** out) {
    if (!out) return FGP_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return FGP_ERR_CUDA;
    fgp_model* m = new (std::nothrow) fgp_model();
    if (!m) return FGP_ERR_CUDA;
    m->device = device;
    DeviceGuard dg(device);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // st2 = panel stream of the look-ahead Cholesky: highest priority
    bool ok = cudaStreamCreateWithPriority(&m->st, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&m->st2, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&m->st3, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evC, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evD, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreate(&m->ev0) == cudaSuccess && cudaEventCreate(&m->ev1) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evA, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evB, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evTop, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evRest, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evCopy[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&m->evCopy[1], cudaEventDisableTiming) == cudaSuccess &&
              potrf_head_prepare() == cudaSuccess &&
              cudaMalloc(&m->info_d, sizeof(int)) == cudaSuccess &&
              cudaMallocHost(&m->info_h, sizeof(int)) == cudaSuccess && potrf_prepare() == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        fgp_destroy(m);
        return FGP_ERR_CUDA;
    }
    *out = m;
    return FGP_OK;
}

FGP_EXPORT int fgp_destroy(fgp_model* m) {
    if (!m) return FGP_OK;
    {
        DeviceGuard dg(m->device);
        if (m->st) cudaStreamSynchronize(m->st);
        if (m->st2) cudaStreamSynchronize(m->st2);
        if (m->st3) cudaStreamSynchronize(m->st3);
        comm_release(m);
        for (DevBuf* b : {&m->xr, &m->xc, &m->nc, &m->nr, &m->cmean, &m->y, &m->z, &m->alpha, &m->work, &m->L, &m->inv,
                          &m->invT, &m->staging, &m->qr, &m->qc, &m->qnc, &m->qnr, &m->bt, &m->partial, &m->mean_d,
                          &m->var_d, &m->scalars, &m->kqq, &m->U, &m->Kinv, &m->lml_partial, &m->lml_rows, &m->Wp, &m->Pscr, &m->pbuf[0],
                          &m->pbuf[1], &m->ozDigits, &m->ozScale, &m->ozL, &m->ozLscale, &m->ozU})
            b->release();
        if (m->head_sync) cudaFree(m->head_sync);
        for (cudaEvent_t e : {m->evTop, m->evRest, m->evCopy[0], m->evCopy[1]})
            if (e) cudaEventDestroy(e);
        if (m->info_d) cudaFree(m->info_d);
        if (m->info_h) cudaFreeHost(m->info_h);
        if (m->pinned) cudaFreeHost(m->pinned);
        if (m->ev0) cudaEventDestroy(m->ev0);
        if (m->ev1) cudaEventDestroy(m->ev1);
        if (m->evA) cudaEventDestroy(m->evA);
        if (m->evB) cudaEventDestroy(m->evB);
        if (m->evC) cudaEventDestroy(m->evC);
        if (m->evD) cudaEventDestroy(m->evD);
        m->prof.destroy();
        if (m->st) cudaStreamDestroy(m->st);
        if (m->st2) cudaStreamDestroy(m->st2);
        if (m->st3) cudaStreamDestroy(m->st3);
    }
    delete m;
    return FGP_OK;
}

FGP_EXPORT const char* fgp_last_error(const fgp_model* m) { return m ? m->err.c_str() : "null handle"; }
FGP_EXPORT int64_t fgp_failed_column(const fgp_model* m) { return m ? m->failed_col : -1; }
FGP_EXPORT int64_t fgp_num_samples(const fgp_model* m) { return m ? m->n : 0; }
FGP_EXPORT int64_t fgp_num_dims(const fgp_model* m) { return m ? m->d : 0; }
FGP_EXPORT double fgp_last_device_ms(const fgp_model* m) { return m ? (double)m->last_ms : 0.0; }
FGP_EXPORT int64_t fgp_last_launch_count(const fgp_model* m) { return m ? m->launches : 0; }

FGP_EXPORT int fgp_set_profiling(fgp_model* m, int on) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    m->profiling = on != 0;
    m->prof.reset();
    return FGP_OK;
}
FGP_EXPORT int fgp_set_option(fgp_model* m, int option, int64_t value) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    switch (option) {
        case FGP_OPT_LOOKAHEAD: m->lookahead = value != 0; return FGP_OK;
        case FGP_OPT_HEAD: m->head_schedule = value != 0; return FGP_OK;
        case FGP_OPT_TCGEN05: m->tcgen05 = value != 0; return FGP_OK;
        case FGP_OPT_SHARD_PIPE: m->shard_pipe = value < 0 ? -1 : (value != 0); return FGP_OK;
        default: return fail(m, FGP_ERR_BAD_ARG, "unknown option");
    }
}
FGP_EXPORT int fgp_profile_summary(const fgp_model* m, double* ms, double* flops, int64_t* count) {
    if (!m || !ms || !flops || !count) return FGP_ERR_BAD_ARG;
    for (int i = 0; i < PROF_NCLASS; ++i) {
        ms[i] = m->prof.ms[i];
        flops[i] = m->prof.flops[i];
        count[i] = m->prof.count[i];
    }
    return FGP_OK;
}

FGP_EXPORT void* fgp_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
FGP_EXPORT void fgp_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

// =================================================================================================================
// fit
namespace {
// EMatrix::new + Input::into_dmatrix (mod.rs:142-148): make the training inputs resident (row-major, padded, centred)
int reserve_inputs(fgp_model* m, int64_t n, int64_t d) {
    if (n <= 0 || d <= 0) return fail(m, FGP_ERR_BAD_ARG, "bad training matrix");
    m->fitted = false;
    const int64_t np = round_up(n, TILE), dp = round_up(d, 4);
    if (np > m->cap || dp != m->dp) {
        const int64_t cap = std::max(np, m->cap);
        FGP_TRY(reserve_training(m, cap, dp, false));
        CU(m, m->L.reserve((size_t)cap * cap));
        m->cap = cap;
    }
    m->n = n;
    m->d = d;
    m->dp = dp;
    m->np = np;
    CU(m, m->staging.reserve((size_t)n * d));
    return FGP_OK;
}
// staging (n x d column-major, ld = n) -> padded row-major raw / centred points and their norms
int convert_staged_inputs(fgp_model* m) {
    const int64_t n = m->n, np = m->np;
    col_mean_kernel<<<(unsigned)m->d, 256, 0, m->st>>>(m->staging.p, n, n, m->cmean.p);
    convert_points_kernel<<<(unsigned)((np + 255) / 256), 256, 0, m->st>>>(m->staging.p, n, n, (int)m->d, (int)m->dp,
                                                                          m->cmean.p, 0, np, m->xr.p, m->xc.p, m->nc.p, m->nr.p);
    m->launches += 2;
    return FGP_OK;
}
int set_inputs(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d) {
    if (!X || n <= 0 || d <= 0 || ldx < n) return fail(m, FGP_ERR_BAD_ARG, "bad training matrix");
    FGP_TRY(reserve_inputs(m, n, d));
    FGP_TRY(upload_colmajor(m, X, ldx, n, d));
    FGP_TRY(convert_staged_inputs(m));
    CU(m, cudaMemsetAsync(m->y.p, 0, m->np * sizeof(double), m->st));
    return FGP_OK;
}
}  // namespace

FGP_EXPORT int fgp_set_inputs(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    begin_timed(m);
    FGP_TRY(set_inputs(m, X, ldx, n, d));
    return end_timed(m);
}

FGP_EXPORT int fgp_fit(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
                       const fgp_kernel_desc* kernel, double noise, int has_eps, double eps) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!y_resid) return fail(m, FGP_ERR_BAD_ARG, "null training outputs");
    if (!(noise >= 0.0)) return fail(m, FGP_ERR_BAD_ARG, "The noise parameter should non-negative");  // mod.rs:150
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    FGP_TRY(set_inputs(m, X, ldx, n, d));
    CU(m, cudaMemcpyAsync(m->y.p, y_resid, n * sizeof(double), cudaMemcpyHostToDevice, m->st));
    int rc = factor_resident(m, kernel, kt, noise, has_eps, eps);
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

FGP_EXPORT int fgp_refit(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int has_eps, double eps) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->n <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no resident training set");
    if (!(noise >= 0.0)) return fail(m, FGP_ERR_BAD_ARG, "The noise parameter should non-negative");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    int rc = factor_resident(m, kernel, kt, noise, has_eps, eps);
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

FGP_EXPORT int fgp_set_outputs(fgp_model* m, const double* y_resid, int64_t n) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->n <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no resident training set");
    if (!y_resid || n != m->n) return fail(m, FGP_ERR_BAD_ARG, "output vector length mismatch");
    begin_timed(m);
    CU(m, cudaMemcpyAsync(m->y.p, y_resid, n * sizeof(double), cudaMemcpyHostToDevice, m->st));
    if (m->fitted) solve_alpha(m);
    return end_timed(m);
}

// =================================================================================================================
// predict
FGP_EXPORT int fgp_predict_mean(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                                double* mean_wo_prior) {
    if (!mean_wo_prior) return m ? fail(m, FGP_ERR_BAD_ARG, "null output") : FGP_ERR_BAD_ARG;
    return predict_common(m, kernel, Xq, ldq, q, mean_wo_prior, nullptr);
}
FGP_EXPORT int fgp_predict_var(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                               double* var) {
    if (!var) return m ? fail(m, FGP_ERR_BAD_ARG, "null output") : FGP_ERR_BAD_ARG;
    return predict_common(m, kernel, Xq, ldq, q, nullptr, var);
}
FGP_EXPORT int fgp_predict_mean_var(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                                    double* mean_wo_prior, double* var) {
    if (!var || !mean_wo_prior) return m ? fail(m, FGP_ERR_BAD_ARG, "null output") : FGP_ERR_BAD_ARG;
    return predict_common(m, kernel, Xq, ldq, q, mean_wo_prior, var);
}

FGP_EXPORT int fgp_stage_queries(fgp_model* m, const double* Xq, int64_t ldq, int64_t q) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    begin_timed(m);
    FGP_TRY(stage_queries(m, Xq, ldq, q));
    return end_timed(m);
}
FGP_EXPORT int fgp_predict_staged(fgp_model* m, const fgp_kernel_desc* kernel, int want_mean, int want_var) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted || m->q <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no staged queries");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    FGP_TRY(predict_device(m, kernel, kt, want_mean, want_var));
    return end_timed(m);
}
FGP_EXPORT int fgp_fetch_predictions(fgp_model* m, double* mean_wo_prior, double* var) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->q <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no staged queries");
    return fetch_predictions(m, mean_wo_prior, var);
}

// predict_covariance (mod.rs:329-350): Kqq - kl^T kl ; sample_at (mod.rs:371-384): Kqq - Knq^T (K^-1 Knq).
// Both are the same matrix; on the device it is formed once as Kqq - Bt Bt^T with Bt = (L^-1 Knq)^T.
FGP_EXPORT int fgp_predict_cov(fgp_model* m, const fgp_kernel_desc* kernel, const double* Xq, int64_t ldq, int64_t q,
                               int mode, double* cov, int64_t ldc, double* mean_wo_prior) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!cov || ldc < q || (mode != 0 && mode != 1)) return fail(m, FGP_ERR_BAD_ARG, "bad covariance output");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    FGP_TRY(stage_queries(m, Xq, ldq, q));
    m->force_batched_predict = true;  // the covariance needs the transposed solve buffer Bt whatever q is
    const int rc_pred = predict_device(m, kernel, kt, mean_wo_prior != nullptr, 1);
    m->force_batched_predict = false;
    FGP_TRY(rc_pred);
    const int64_t qp = m->qp;
    CU(m, m->kqq.reserve((size_t)qp * qp));
    PairArgs pa = query_pair_args(m);
    pa.xb_c = m->qc.p;
    pa.xb_r = m->qr.p;
    pa.nb = m->qnc.p;
    pa.cols = qp;
    write_covariance(m, kt, kernel, pa, m->kqq.p, qp, q, q, 0.0);
    GemmArgs g{};
    g.C = m->kqq.p; g.ldc = qp;
    g.A = m->bt.p; g.lda = qp;
    g.B = m->bt.p; g.ldb = qp;
    g.M = g.N = (int)qp; g.K = (int)m->np;
    g.alpha = -1.0; g.beta_one = 1; g.lower = 0; g.k_from_tile = 0;
    m->launches += gemm_nt_launch(g, m->ctx()) > 0;
    FGP_TRY(end_timed(m));
    // q x q result straight into the caller's buffer
    CU(m, cudaMemcpy2DAsync(cov, ldc * sizeof(double), m->kqq.p, qp * sizeof(double), q * sizeof(double), q,
                            cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    if (mean_wo_prior) return fetch_predictions(m, mean_wo_prior, nullptr);
    return FGP_OK;
}

// =================================================================================================================
// model selection
FGP_EXPORT int fgp_likelihood(fgp_model* m, const fgp_kernel_desc* kernel, double noise, double* out) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!out) return fail(m, FGP_ERR_BAD_ARG, "null output");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    CU(m, m->scalars.reserve(8));
    const DevKernel dk = to_dev(kernel);
    reduce_kernel<0><<<1, 256, 0, m->st>>>(m->z.p, nullptr, m->n, dk, 0.0, m->scalars.p);                 // ||L^-1 y||^2
    reduce_kernel<1><<<1, 256, 0, m->st>>>(m->nr.p, nullptr, m->n, dk, noise * noise, m->scalars.p + 1);  // mod.rs:208-213
    m->launches += 2;
    FGP_TRY(ensure_pinned(m, 8));
    CU(m, cudaMemcpyAsync(m->pinned, m->scalars.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    FGP_TRY(end_timed(m));
    const double data_fit = m->pinned[0], penalty = m->pinned[1];
    const double norm = (double)m->n * std::log(2.0 * M_PI);
    *out = -(data_fit + penalty + norm) / 2.0;  // mod.rs:216-219
    return FGP_OK;
}

FGP_EXPORT int fgp_lml_gradient(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int scaled, double* scale_out,
                                double* grads) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!grads) return fail(m, FGP_ERR_BAD_ARG, "null output");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    int rc = lml_gradient_device(m, kernel, kt, noise, scaled, scale_out, grads);
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

// The same gradient, collectively: every rank of the communicator calls it after fgp_fit_sharded / fgp_refit_sharded (all
// hold the full factor); the O(n^3) inverse is split over the ranks (csrc/lml.cu) and every rank returns the same values.
FGP_EXPORT int fgp_lml_gradient_sharded(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int scaled, double* scale_out,
                                        double* grads) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->comm) return fail(m, FGP_ERR_COMM, "fgp_comm_init_rank has not been called");
    KernelTraits kt;
    begin_timed(m);
    const auto local = [&]() -> int {
        if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
        if (!grads) return fail(m, FGP_ERR_BAD_ARG, "null output");
        return check_kernel(m, kernel, &kt);
    };
    FGP_TRY(comm_agree(m, local()));
    int rc = lml_gradient_sharded_device(m, kernel, kt, noise, scaled, scale_out, grads);
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

FGP_EXPORT int fgp_mean_pair_distance(fgp_model* m, double* out) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->n <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no resident training set");
    if (!out) return fail(m, FGP_ERR_BAD_ARG, "null output");
    begin_timed(m);
    int rc = mean_pair_distance_device(m, out);
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

FGP_EXPORT int fgp_linear_prior_fit(fgp_model* m, const double* y, double* weights, double* intercept) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (m->n <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no resident training set");
    if (!y || !weights || !intercept) return fail(m, FGP_ERR_BAD_ARG, "null argument");
    begin_timed(m);
    int rc = linear_prior_fit_device(m, y, weights, intercept);
    if (rc == FGP_ERR_BAD_ARG) rc = fail(m, FGP_ERR_BAD_ARG, "linear prior fit supports up to 44 input dimensions");
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

// =================================================================================================================
// state transfer
FGP_EXPORT int fgp_download_factor(fgp_model* m, double* L, int64_t ldl) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!L || ldl < m->n) return fail(m, FGP_ERR_BAD_ARG, "bad factor buffer");
    const int64_t n = m->n;
    CU(m, cudaMemcpy2DAsync(L, ldl * sizeof(double), m->L.p, m->cap * sizeof(double), n * sizeof(double), n,
                            cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    const double nanv = std::numeric_limits<double>::quiet_NaN();  // algebra/mod.rs:67: the upper triangle is never written
    for (int64_t c = 1; c < n; ++c)
        for (int64_t r = 0; r < c; ++r) L[r + c * ldl] = nanv;
    return FGP_OK;
}

FGP_EXPORT int fgp_download_alpha(fgp_model* m, double* alpha) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!alpha) return fail(m, FGP_ERR_BAD_ARG, "null output");
    CU(m, cudaMemcpyAsync(alpha, m->alpha.p, m->n * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    return FGP_OK;
}

namespace {
// strict upper triangle of the n_pad x n_pad matrix <- 0 (a serialised nalgebra factor carries NaN there, algebra/mod.rs:67)
__global__ void zero_strict_upper_kernel(double* A, int64_t ld, int64_t np) {
    const int64_t c = blockIdx.x;
    for (int64_t r = threadIdx.x; r < c; r += blockDim.x) A[r + c * ld] = 0.0;
}
// out[r + i*ldo] = Kinv[r, cols[i]] from the lower triangle of the symmetric inverse
__global__ void gather_sym_columns_kernel(const double* __restrict__ Kinv, int64_t ld, int64_t n, const int64_t* __restrict__ cols,
                                          double* __restrict__ out, int64_t ldo) {
    const int64_t c = cols[blockIdx.y];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r + (int64_t)blockIdx.y * ldo] = (r >= c) ? Kinv[r + c * ld] : Kinv[c + r * ld];
}
}  // namespace

namespace {
// Deterministic digest of the lower triangle of the factor: per column c (one CTA, fixed-order tree), three sums
//   d0 = sum L[r,c]      d1 = sum L[r,c]^2      d2 = sum L[r,c] * w(r,c),  w = ((31 r + 17 c) mod 1009) + 1
// then one thread folds the columns in order.  Bitwise-equal factors give bitwise-equal digests (and a single differing
// element changes d2 with overwhelming probability): cross-rank / sharded-vs-single comparisons without moving 8 n^2 bytes.
__global__ void __launch_bounds__(256) digest_columns_kernel(const double* __restrict__ L, int64_t ld, int64_t n, double* partial) {
    __shared__ double red[3][256];
    const int64_t c = blockIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t r = c + threadIdx.x; r < n; r += 256) {
        const double v = L[r + c * ld];
        s0 += v;
        s1 = fma(v, v, s1);
        s2 = fma(v, (double)((31 * r + 17 * c) % 1009 + 1), s2);
    }
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[c * 3 + threadIdx.x] = red[threadIdx.x][0];
}
__global__ void digest_final_kernel(const double* __restrict__ partial, int64_t n, double* out) {
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int64_t c = 0; c < n; ++c) s += partial[c * 3 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}
}  // namespace

// out[0..2] = (sum, sum of squares, position-weighted sum) over the lower triangle of the resident factor; deterministic.
FGP_EXPORT int fgp_factor_digest(fgp_model* m, double* out) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!out) return fail(m, FGP_ERR_BAD_ARG, "null output");
    CU(m, m->staging.reserve((size_t)3 * m->n + 8));
    digest_columns_kernel<<<(unsigned)m->n, 256, 0, m->st>>>(m->L.p, m->cap, m->n, m->staging.p);
    digest_final_kernel<<<1, 32, 0, m->st>>>(m->staging.p, m->n, m->staging.p + 3 * m->n);
    FGP_TRY(ensure_pinned(m, 8));
    CU(m, cudaMemcpyAsync(m->pinned, m->staging.p + 3 * m->n, 3 * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    CU(m, cudaGetLastError());
    for (int i = 0; i < 3; ++i) out[i] = m->pinned[i];
    return FGP_OK;
}

// Restore a handle from a serialised model (serde round trip of GaussianProcess, mod.rs:58; EMatrix / EVector
// extendable_matrix.rs:14,62; nalgebra's Cholesky keeps the full n x n matrix with the factor in its lower triangle):
// training inputs, residual outputs and the factor go back to the device WITHOUT refitting; what the device path caches on
// top of the reference's state (inverse diagonal blocks, alpha = K^-1 y, z = L^-1 y) is rebuilt from L.
FGP_EXPORT int fgp_upload_state(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
                                const double* L, int64_t ldl) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!y_resid || !L || ldl < n) return fail(m, FGP_ERR_BAD_ARG, "bad serialised state");
    begin_timed(m);
    FGP_TRY(set_inputs(m, X, ldx, n, d));
    CU(m, cudaMemcpyAsync(m->y.p, y_resid, n * sizeof(double), cudaMemcpyHostToDevice, m->st));
    const int64_t np = m->np;
    CU(m, cudaMemset2DAsync(m->L.p, m->cap * sizeof(double), 0, np * sizeof(double), np, m->st));
    launch_set_identity(m->L.p, m->cap, np, m->st);  // padding block = I (overwritten on the first n diagonal entries)
    CU(m, cudaMemcpy2DAsync(m->L.p, m->cap * sizeof(double), L, ldl * sizeof(double), n * sizeof(double), n,
                            cudaMemcpyHostToDevice, m->st));
    zero_strict_upper_kernel<<<(unsigned)np, 256, 0, m->st>>>(m->L.p, m->cap, np);
    launch_diag_inverse(m->L.p, m->cap, np / TILE, m->inv.p, m->invT.p, m->st);
    m->launches += 3;
    FGP_TRY(rebuild_panel_inverses(m));
    solve_alpha(m);
    m->failed_col = -1;
    m->fitted = true;
    m->kinv_valid = false;
    return end_timed(m);
}

// Selected columns of K^-1 = covmat_cholesky.inverse() (optimizer.rs:32, :169) as left on the device by the last
// fgp_lml_gradient call (lower triangle stored; symmetrised here).  out is n x ncols column-major.  Tests / diagnostics.
FGP_EXPORT int fgp_inverse_columns(fgp_model* m, const int64_t* cols, int64_t ncols, double* out, int64_t ldo) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted || !m->kinv_valid) return fail(m, FGP_ERR_NOT_FITTED, "no inverse on the device: call fgp_lml_gradient first");
    if (!cols || !out || ncols <= 0 || ncols > 4096 || ldo < m->n) return fail(m, FGP_ERR_BAD_ARG, "bad column selection");
    for (int64_t i = 0; i < ncols; ++i)
        if (cols[i] < 0 || cols[i] >= m->n) return fail(m, FGP_ERR_BAD_ARG, "column out of range");
    const int64_t n = m->n;
    CU(m, m->staging.reserve((size_t)n * ncols + (size_t)ncols));
    int64_t* dcols = reinterpret_cast<int64_t*>(m->staging.p + (size_t)n * ncols);
    CU(m, cudaMemcpyAsync(dcols, cols, ncols * sizeof(int64_t), cudaMemcpyHostToDevice, m->st));
    gather_sym_columns_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)ncols), 256, 0, m->st>>>(m->Kinv.p, m->np, n, dcols,
                                                                                                 m->staging.p, n);
    CU(m, cudaMemcpy2DAsync(out, ldo * sizeof(double), m->staging.p, n * sizeof(double), n * sizeof(double), ncols,
                            cudaMemcpyDeviceToHost, m->st));
    CU(m, cudaStreamSynchronize(m->st));
    CU(m, cudaGetLastError());
    return FGP_OK;
}

// =================================================================================================================
// add_samples (mod.rs:173-190 + algebra/mod.rs:97-126)
FGP_EXPORT int fgp_add_samples(fgp_model* m, const double* Xnew, int64_t ldx, int64_t k, const double* ynew_resid,
                               const fgp_kernel_desc* kernel, double noise, int has_eps, double eps) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->fitted) return fail(m, FGP_ERR_NOT_FITTED, "model is not fitted");
    if (!Xnew || !ynew_resid || k <= 0 || ldx < k) return fail(m, FGP_ERR_BAD_ARG, "bad sample matrix");
    if (!(noise >= 0.0)) return fail(m, FGP_ERR_BAD_ARG, "The noise parameter should non-negative");
    KernelTraits kt;
    FGP_TRY(check_kernel(m, kernel, &kt));
    begin_timed(m);
    const int64_t n_old = m->n, n_new = n_old + k, np_new = round_up(n_new, TILE);
    // --- capacity (EMatrix::add_rows grows by x1.5, extendable_matrix.rs:30-49) -------------------------------
    if (np_new > m->cap) {
        const int64_t cap_new = std::max(np_new, round_up(m->cap + m->cap / 2, TILE));
        FGP_TRY(reserve_training(m, cap_new, m->dp, true));
        DevBuf Lnew;
        CU(m, Lnew.reserve((size_t)cap_new * cap_new));
        CU(m, cudaMemcpy2DAsync(Lnew.p, cap_new * sizeof(double), m->L.p, m->cap * sizeof(double), m->np * sizeof(double),
                                m->np, cudaMemcpyDeviceToDevice, m->st));
        CU(m, cudaStreamSynchronize(m->st));
        m->L.release();
        m->L = Lnew;
        m->cap = cap_new;
    }
    // --- new points: same centring as the resident ones (cmean is only a numerical shift, any value is valid) ---
    FGP_TRY(upload_colmajor(m, Xnew, ldx, k, m->d));
    convert_points_kernel<<<(unsigned)((np_new - n_old + 255) / 256), 256, 0, m->st>>>(
        m->staging.p, k, k, (int)m->d, (int)m->dp, m->cmean.p, n_old, np_new - n_old, m->xr.p, m->xc.p, m->nc.p, m->nr.p);
    m->launches += 1;
    CU(m, cudaMemsetAsync(m->y.p + n_old, 0, (np_new - n_old) * sizeof(double), m->st));
    CU(m, cudaMemcpyAsync(m->y.p + n_old, ynew_resid, k * sizeof(double), cudaMemcpyHostToDevice, m->st));
    // --- block rows >= jb are (re)built: Gram rows, forward solve against the frozen block columns, trailing potrf ---
    const int64_t jb = n_old / TILE;  // first block column touched by the new rows
    m->n = n_new;
    m->np = np_new;
    CU(m, cudaMemsetAsync(m->info_d, 0, sizeof(int), m->st));
    PairArgs pa = train_pair_args(m);
    pa.row_tile0 = (int)jb;
    write_covariance(m, kt, kernel, pa, m->L.p, m->cap, n_new, n_new, noise * noise);
    // rows [jb*128, np_new) x columns [0, jb*128):  A <- A * L11^-T, and the trailing block gets -= A A^T
    double* Arows = m->L.p + jb * TILE;  // row offset inside every column
    PotrfCounters cnt;
    const PotrfLookahead la{m->st2, m->evA, m->evB, m->st3, m->evC, m->evD};
    const int64_t Mrows = np_new - jb * TILE;
    if (m->head_schedule) {
        const bool had_w = m->w_valid;
        // the digit slices the last full fit kept still mirror L11 (old rows x old panels): the solve of the new rows against the
        // old columns may use them; afterwards they are stale (they do not cover the new rows)
        const bool oz_ok = m->tcgen05 && m->ozL_valid && m->ozOffBytes.size() >= (size_t)m->pstart.size();
        m->ozL_valid = false;
        PotrfWork w;
        int64_t p0 = 0;
        FGP_TRY(prepare_head_work(m, jb, &w, &p0));  // panel table: the old panels (the last one cut at jb), then the new ones
        if (had_w && p0 > 0 && Mrows <= m->cap) {
            OzPanelStore ozs{};
            const bool use_oz = oz_ok && (int64_t)m->ozOffBytes.size() >= p0 && m->ozDigits.p;
            if (use_oz)
                ozs = OzPanelStore{reinterpret_cast<const int8_t*>(m->ozL.p), m->ozLscale.p, m->ozOffBytes.data(), m->ozOffRows.data(),
                                   reinterpret_cast<int8_t*>(m->ozDigits.p), m->ozScale.p};
            m->launches += trsm_fwd_t_panels(Arows, m->cap, Mrows, m->L.p, m->cap, m->Wp.p, m->pstart.data(), p0, jb,
                                             np_new / TILE, m->pbuf[0].p, m->pbuf[1].p, m->ctx(), m->lookahead ? m->st2 : nullptr,
                                             m->evA, m->evB, m->evC, use_oz ? &ozs : nullptr);
        } else
            m->launches += trsm_fwd_t(Arows, m->cap, Mrows, m->L.p, m->cap, m->inv.p, 0, jb, Arows + jb * TILE * m->cap, m->ctx());
        potrf_lower_head(m->L.p, m->cap, np_new, jb, w, p0, has_eps, eps, m->info_d, m->ctx(), m->lookahead ? &la : nullptr, &cnt);
        m->w_valid = had_w;
    } else {
        m->launches += trsm_fwd_t(Arows, m->cap, Mrows, m->L.p, m->cap, m->inv.p, 0, jb, Arows + jb * TILE * m->cap, m->ctx());
        potrf_lower(m->L.p, m->cap, np_new, jb, m->inv.p, m->invT.p, has_eps, eps, m->info_d, m->ctx(),
                    m->lookahead ? &la : nullptr, &cnt);
        m->pstart.clear();
        m->w_valid = false;
    }
    m->launches += cnt.launches;
    CU(m, cudaMemcpyAsync(m->info_h, m->info_d, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    solve_alpha(m);
    int rc2 = end_timed(m);
    if (rc2 == FGP_OK && *m->info_h == HEAD_TIMEOUT) rc2 = fail(m, FGP_ERR_CUDA, "internal error: a dependency wait inside potrf_head_kernel timed out");
    if (rc2 != FGP_OK || *m->info_h != 0) {
        // The block rows from jb on (old rows of the last partial tile included) have been rebuilt and are now garbage: the
        // handle goes back to the OLD sample count and needs a full refit (fgp_refit / fgp_fit), which the resident inputs
        // and outputs of the first n_old samples still allow; the host twin (gp.py) has not appended the new rows either.
        m->n = n_old;
        m->np = round_up(n_old, TILE);
        m->fitted = false;
        if (rc2 != FGP_OK) return rc2;
        m->failed_col = *m->info_h - 1;
        return fail(m, FGP_ERR_NOT_POSDEF, "Cholesky update failed at column " + std::to_string(m->failed_col));
    }
    m->failed_col = -1;
    m->kinv_valid = false;
    return FGP_OK;
}

// =================================================================================================================
// Cholesky of a caller-supplied SPD matrix with the same device factorisation (MultivariateNormal::new,
// multivariate_normal.rs:54-59: `covariance.cholesky().expect(..).unpack()`): A (n x n column-major, lower triangle
// read) is overwritten by L with the strict upper triangle zeroed (unpack()).
FGP_EXPORT int fgp_cholesky_lower(int device, double* A, int64_t lda, int64_t n, int64_t* failed_col) {
    if (!A || n <= 0 || lda < n) return FGP_ERR_BAD_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return FGP_ERR_CUDA;
    DeviceGuard dg(device);
    if (potrf_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    const int64_t np = round_up(n, TILE);
    double *dA = nullptr, *dinv = nullptr;
    int* dinfo = nullptr;
    int info = 0;
    int rc = FGP_OK;
    if (cudaMalloc(&dA, (size_t)np * np * 8) != cudaSuccess || cudaMalloc(&dinv, (size_t)2 * np * TILE * 8) != cudaSuccess ||
        cudaMalloc(&dinfo, sizeof(int)) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        cudaMemsetAsync(dA, 0, (size_t)np * np * 8, 0);
        cudaMemsetAsync(dinfo, 0, sizeof(int), 0);
        launch_set_identity(dA, np, np, 0);  // padding block = I
        cudaMemcpy2DAsync(dA, (size_t)np * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyHostToDevice, 0);
        PotrfCounters cnt;
        potrf_lower(dA, np, np, 0, dinv, dinv + np * TILE, 0, 0.0, dinfo, LaunchCtx{}, nullptr, &cnt);
        cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, 0);
        cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)np * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost, 0);
        if (cudaStreamSynchronize(0) != cudaSuccess) rc = FGP_ERR_CUDA;
    }
    cudaFree(dA);
    cudaFree(dinv);
    cudaFree(dinfo);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    if (rc != FGP_OK) return rc;
    if (failed_col) *failed_col = info ? info - 1 : -1;
    if (info) return FGP_ERR_NOT_POSDEF;
    for (int64_t c = 1; c < n; ++c)
        for (int64_t r = 0; r < c; ++r) A[r + c * lda] = 0.0;
    return FGP_OK;
}

// =================================================================================================================
// multi-GPU: one process per GPU, NCCL panel broadcasts (sharded.cu)
FGP_EXPORT int fgp_comm_unique_id(void* id_out, size_t bytes) {
    if (!id_out || bytes < sizeof(ncclUniqueId)) return FGP_ERR_BAD_ARG;
    ncclUniqueId id;
    const NcclApi* nccl = nccl_api();
    if (!nccl || nccl->GetUniqueId(&id) != ncclSuccess) return FGP_ERR_COMM;
    std::memset(id_out, 0, bytes);
    std::memcpy(id_out, &id, sizeof(id));
    return FGP_OK;
}

FGP_EXPORT int fgp_comm_init_rank(fgp_model* m, const void* id, size_t bytes, int nranks, int rank) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!id || bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
        return fail(m, FGP_ERR_BAD_ARG, "bad communicator arguments");
    comm_release(m);
    fgp_comm* c = new (std::nothrow) fgp_comm();
    if (!c) return fail(m, FGP_ERR_CUDA, "out of host memory");
    c->nranks = nranks;
    c->rank = rank;
    m->comm = c;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->st_comm, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->st_copy, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_col, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_bcast, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_trail[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_trail[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_copy[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_copy[1], cudaEventDisableTiming) != cudaSuccess) {
        comm_release(m);
        return fail(m, FGP_ERR_CUDA, "event creation failed");
    }
    if (nranks > 1) {
        const NcclApi* nccl = nccl_api();
        if (!nccl) {
            comm_release(m);
            return fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
        }
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        ncclResult_t r = nccl->CommInitRank(&c->comm, nranks, uid, rank);
        if (r != ncclSuccess) {
            std::string msg = std::string("ncclCommInitRank: ") + nccl->GetErrorString(r);
            comm_release(m);
            return fail(m, FGP_ERR_COMM, msg);
        }
    }
    return FGP_OK;
}

FGP_EXPORT int fgp_comm_destroy(fgp_model* m) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    comm_release(m);
    return FGP_OK;
}

FGP_EXPORT double fgp_comm_last_bytes(const fgp_model* m) { return (m && m->comm) ? m->comm->bcast_bytes : 0.0; }

// Host-only description of the block-cyclic plan (no GPU needed): panel width in columns, number of panels, how many this
// rank owns and its share of the trailing-update flops.
FGP_EXPORT int fgp_shard_plan(int64_t n, int nranks, int rank, int64_t* panel_cols, int64_t* n_panels, int64_t* n_owned,
                              double* flop_share) {
    if (n <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return FGP_ERR_BAD_ARG;
    // the head schedule (default) uses 512-column panels whatever the size and the number of ranks
    const int64_t np = round_up(n, TILE), nb = np / TILE, PANEL_TILES = HEAD_PANEL / TILE, NP = (nb + PANEL_TILES - 1) / PANEL_TILES;
    int64_t owned = 0;
    double mine = 0.0, total = 0.0;
    for (int64_t p = 0; p < NP; ++p) {
        const double rows = (double)(np - p * PANEL_TILES * TILE);
        const double w = (double)std::min<int64_t>(PANEL_TILES * TILE, np - p * PANEL_TILES * TILE);
        const double f = rows * w * (double)(p * PANEL_TILES * TILE);  // updates this panel receives from the p panels before it
        total += f;
        if (shard_owner(p, nranks) == rank) {
            owned += 1;
            mine += f;
        }
    }
    if (panel_cols) *panel_cols = PANEL_TILES * TILE;
    if (n_panels) *n_panels = NP;
    if (n_owned) *n_owned = owned;
    if (flop_share) *flop_share = total > 0.0 ? mine / total : (rank == 0 ? 1.0 : 0.0);
    return FGP_OK;
}

namespace {
int run_factor_sharded(fgp_model* m, const fgp_kernel_desc* kernel, const KernelTraits& kt, double noise, int has_eps, double eps) {
    if (!m->head_schedule) {
        m->pstart.clear();
        m->w_valid = false;
        return factor_sharded(m, kernel, kt, noise, has_eps, eps);
    }
    PotrfWork w;
    int64_t p0 = 0;
    FGP_TRY(prepare_head_work(m, 0, &w, &p0));
    m->w_valid = false;
    static const bool pipe_env = !(getenv("FGP_SHARD_PIPE") && atoi(getenv("FGP_SHARD_PIPE")) == 0);
    // automatic: row pieces from 3 ranks up.  Measured on C4 (n = 32768): 2 GPUs 115.3 ms in one piece vs 117.6 - 119.2 ms in pieces
    // (work-bound: the chain is hidden either way and the pieces cost launches), 4 GPUs 70.6 vs 70.3 ms, 8 GPUs 58.0 vs 47.0 ms
    const bool pipe = pipe_env && (m->shard_pipe < 0 ? m->comm->nranks >= 3 : m->shard_pipe != 0);
    const int rc = pipe ? factor_sharded_pipe(m, kernel, kt, noise, has_eps, eps, w) : factor_sharded_head(m, kernel, kt, noise, has_eps, eps, w);
    if (rc == FGP_OK) {
        // every rank ends with the full factor, every panel's inverse diagonal block W_p (broadcast beside the panel) and the digit
        // slices of every panel: predict / likelihood / LML gradient run locally like after a single-GPU fit
        m->w_valid = true;
        m->ozL_valid = w.oz_off_bytes != nullptr;
    }
    return rc;
}

int finish_sharded(fgp_model* m, int rc) {
    if (rc != FGP_OK) return rc;
    CU(m, cudaMemcpyAsync(m->info_h, m->info_d, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    solve_alpha(m);
    CU(m, cudaStreamSynchronize(m->st));
    CU(m, cudaGetLastError());
    if (*m->info_h != 0) {
        m->failed_col = *m->info_h - 1;
        m->fitted = false;
        return fail(m, FGP_ERR_NOT_POSDEF, "Cholesky decomposition failed at column " + std::to_string(m->failed_col));
    }
    m->failed_col = -1;
    m->fitted = true;
    m->kinv_valid = false;
    return FGP_OK;
}
// A rank that fails BEFORE the collective part (argument check, allocation, upload) must not leave the others waiting in
// ncclBroadcast: every rank reports its local status here and all of them return an error when any one failed.
int comm_agree(fgp_model* m, int local_rc) {
    fgp_comm* c = m->comm;
    if (!c || c->nranks == 1) return local_rc;
    const NcclApi* nccl = nccl_api();
    if (!nccl) return local_rc != FGP_OK ? local_rc : fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
    *m->info_h = (local_rc != FGP_OK) ? 1 : 0;
    const std::string local_msg = m->err;
    bool ok = cudaMemcpyAsync(m->info_d, m->info_h, sizeof(int), cudaMemcpyHostToDevice, m->st) == cudaSuccess &&
              nccl->AllReduce(m->info_d, m->info_d, 1, ncclInt, ncclMax, c->comm, m->st) == ncclSuccess &&
              cudaMemcpyAsync(m->info_h, m->info_d, sizeof(int), cudaMemcpyDeviceToHost, m->st) == cudaSuccess &&
              cudaStreamSynchronize(m->st) == cudaSuccess;
    if (local_rc != FGP_OK) {
        m->err = local_msg;
        return local_rc;
    }
    if (!ok) return fail(m, FGP_ERR_COMM, "status exchange between the ranks failed");
    if (*m->info_h != 0) return fail(m, FGP_ERR_COMM, "another rank failed before the collective factorisation started");
    return FGP_OK;
}
}  // namespace

// Same contract as fgp_fit, called by EVERY rank of the communicator. X / y_resid are read on rank 0 only (other ranks may
// pass NULL) and reach the other GPUs by ncclBroadcast; every rank ends with the complete factor and alpha.
FGP_EXPORT int fgp_fit_sharded(fgp_model* m, const double* X, int64_t ldx, int64_t n, int64_t d, const double* y_resid,
                               const fgp_kernel_desc* kernel, double noise, int has_eps, double eps) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->comm) return fail(m, FGP_ERR_COMM, "fgp_comm_init_rank has not been called");
    const bool root = m->comm->rank == 0;
    KernelTraits kt;
    begin_timed(m);
    // rank-local part: anything that can fail here is agreed on before the first collective
    const auto local = [&]() -> int {
        if (root && (!X || !y_resid || ldx < n)) return fail(m, FGP_ERR_BAD_ARG, "rank 0 must supply X and y_resid");
        if (!(noise >= 0.0)) return fail(m, FGP_ERR_BAD_ARG, "The noise parameter should non-negative");
        FGP_TRY(check_kernel(m, kernel, &kt));
        FGP_TRY(reserve_inputs(m, n, d));
        FGP_TRY(reserve_sharded(m));
        CU(m, cudaMemsetAsync(m->y.p, 0, m->np * sizeof(double), m->st));
        if (root) {
            FGP_TRY(upload_colmajor(m, X, ldx, n, d));
            CU(m, cudaMemcpyAsync(m->y.p, y_resid, n * sizeof(double), cudaMemcpyHostToDevice, m->st));
        }
        return FGP_OK;
    };
    FGP_TRY(comm_agree(m, local()));
    if (m->comm->nranks > 1) {
        const NcclApi* nccl = nccl_api();
        if (!nccl || nccl->Broadcast(m->staging.p, m->staging.p, (size_t)n * d, ncclDouble, 0, m->comm->comm, m->st) != ncclSuccess ||
            nccl->Broadcast(m->y.p, m->y.p, (size_t)n, ncclDouble, 0, m->comm->comm, m->st) != ncclSuccess)
            return fail(m, FGP_ERR_COMM, "ncclBroadcast of the training set failed");
    }
    FGP_TRY(convert_staged_inputs(m));
    int rc = finish_sharded(m, run_factor_sharded(m, kernel, kt, noise, has_eps, eps));
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

FGP_EXPORT int fgp_refit_sharded(fgp_model* m, const fgp_kernel_desc* kernel, double noise, int has_eps, double eps) {
    if (!m) return FGP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard dg(m->device);
    if (!m->comm) return fail(m, FGP_ERR_COMM, "fgp_comm_init_rank has not been called");
    KernelTraits kt;
    begin_timed(m);
    const auto local = [&]() -> int {
        if (m->n <= 0) return fail(m, FGP_ERR_NOT_FITTED, "no resident training set");
        if (!(noise >= 0.0)) return fail(m, FGP_ERR_BAD_ARG, "The noise parameter should non-negative");
        FGP_TRY(check_kernel(m, kernel, &kt));
        return reserve_sharded(m);
    };
    FGP_TRY(comm_agree(m, local()));
    int rc = finish_sharded(m, run_factor_sharded(m, kernel, kt, noise, has_eps, eps));
    int rc2 = end_timed(m);
    return rc != FGP_OK ? rc : rc2;
}

// =================================================================================================================
// test hook (host only, no GPU): the block -> tile map of the lower-mode GEMM launches, see gemm_tile_decode
FGP_EXPORT int64_t fgp_dbg_lower_tiles_skip(int M, int N, int grp, int stride, int row_skip, int* ti_out, int* tj_out,
                                            int64_t capacity) {
    GemmArgs g{};
    g.M = M; g.N = N; g.K = GEMM_KC; g.lower = 1; g.grp = grp; g.stride = stride; g.row_skip = row_skip;
    const int64_t tiles = gemm_nt_tiles(g);
    gemm_nt_plan(g);
    for (int64_t b = 0; b < tiles && b < capacity; ++b) gemm_tile_decode(g, (int)b, ti_out[b], tj_out[b]);
    return tiles;
}
FGP_EXPORT int64_t fgp_dbg_lower_tiles(int M, int N, int grp, int stride, int* ti_out, int* tj_out, int64_t capacity) {
    return fgp_dbg_lower_tiles_skip(M, N, grp, stride, 0, ti_out, tj_out, capacity);
}

// =================================================================================================================
// test hook: the panel head kernel alone on a host matrix.  A: (128 nt)^2 column-major SPD (lower read) -> L in place (lower);
// W: (128 nt)^2 (ld = 128 nt) <- L^-1 (lower block triangle; strict upper blocks zeroed).  *info_out = the kernel's info word.
FGP_EXPORT int fgp_dbg_potrf_head(int device, double* A, int nt, double* W, int has_sub, double sub, int* info_out, int reps,
                                  double* ms_out) {
    if (!A || !W || nt < 1 || nt > HEAD_PANEL / TILE) return FGP_ERR_BAD_ARG;
    DeviceGuard dg(device);
    if (potrf_head_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    const int64_t n = (int64_t)nt * TILE;
    double *dA = nullptr, *dinv = nullptr, *dW = nullptr, *dP = nullptr;
    int *dsync = nullptr, *dinfo = nullptr;
    int rc = FGP_OK;
    if (cudaMalloc(&dA, n * n * 8) != cudaSuccess || cudaMalloc(&dinv, n * TILE * 8) != cudaSuccess ||
        cudaMalloc(&dW, HEAD_PANEL * HEAD_PANEL * 8) != cudaSuccess || cudaMalloc(&dP, HEAD_PANEL * HEAD_PANEL * 8) != cudaSuccess ||
        cudaMalloc(&dsync, HEAD_SYNC_INTS * 4) != cudaSuccess || cudaMalloc(&dinfo, 4) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        cudaMemcpy(dA, A, n * n * 8, cudaMemcpyHostToDevice);
        cudaMemset(dW, 0, HEAD_PANEL * HEAD_PANEL * 8);
        cudaMemset(dsync, 0, HEAD_SYNC_INTS * 4);
        cudaMemset(dinfo, 0, 4);
        launch_potrf_head(dA, n, nt, dinv, dW, dP, dsync, has_sub, sub, dinfo, 0, LaunchCtx{});
        if (cudaDeviceSynchronize() != cudaSuccess) rc = FGP_ERR_CUDA;
        if (ms_out && reps > 0 && rc == FGP_OK) {  // timing: the same launch on a fresh copy of A, CUDA events around the kernel only
            double* dA2 = nullptr;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            if (cudaMalloc(&dA2, n * n * 8) == cudaSuccess) {
                float total = 0.f;
                for (int r = 0; r < reps; ++r) {
                    cudaMemcpy(dA2, A, n * n * 8, cudaMemcpyHostToDevice);
                    cudaMemset(dsync, 0, HEAD_SYNC_INTS * 4);
                    cudaEventRecord(e0, nullptr);
                    launch_potrf_head(dA2, n, nt, dinv, dW, dP, dsync, has_sub, sub, dinfo, 0, LaunchCtx{});
                    cudaEventRecord(e1, nullptr);
                    cudaEventSynchronize(e1);
                    float t = 0.f;
                    cudaEventElapsedTime(&t, e0, e1);
                    total += t;
                }
                *ms_out = total / reps;
                cudaFree(dA2);
            }
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        cudaMemcpy(A, dA, n * n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy2D(W, n * 8, dW, HEAD_PANEL * 8, n * 8, n, cudaMemcpyDeviceToHost);
        if (info_out) cudaMemcpy(info_out, dinfo, 4, cudaMemcpyDeviceToHost);
    }
    for (void* q : {(void*)dA, (void*)dinv, (void*)dW, (void*)dP, (void*)dsync, (void*)dinfo}) cudaFree(q);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    return rc;
}

// =================================================================================================================
// test hook: the production GEMM on host matrices
// test hook, host only: the row pieces of a panel with `below` rows under its diagonal block (sharded.cuh shard_pieces)
FGP_EXPORT int fgp_dbg_shard_pieces(int64_t below, int64_t pipe_rows, int64_t* first_row, int64_t* height, int cap) {
    if (below < 1024 || below % 128 != 0 || pipe_rows < 128 || !first_row || !height) return -1;
    std::vector<std::pair<int64_t, int64_t>> rh;
    shard_pieces(below, pipe_rows, rh);
    if ((int)rh.size() > cap) return -1;
    for (size_t i = 0; i < rh.size(); ++i) {
        first_row[i] = rh[i].first;
        height[i] = rh[i].second;
    }
    return (int)rh.size();
}
FGP_EXPORT double fgp_dbg_exp(double x) { return exp_nonpos(x); }
// host twin of the table-assisted exp of the pair-tile fast path (unit scale)
FGP_EXPORT double fgp_dbg_exp_tab(double x) {
    static const double tab[256] = {FGP_EXP2_TABLE_256};
    return exp_nonpos_tab(x, tab);
}

FGP_EXPORT int fgp_dbg_gemm_occupancy(int device) {
    DeviceGuard dg(device);
    return gemm_nt_occupancy(64);
}
FGP_EXPORT int fgp_dbg_gemm_cta_rows(int M, int N, int lower, int num_sms) {
    GemmArgs g{};
    g.M = M; g.N = N; g.K = GEMM_KC; g.lower = lower;
    return gemm_nt_cta_rows(gemm_nt_tiles(g), num_sms);
}
FGP_EXPORT int fgp_dbg_gemm_occupancy32(int device) {
    DeviceGuard dg(device);
    return gemm_nt_occupancy(32);
}

namespace {
__global__ void fill_random_kernel(double* p, int64_t n, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);  // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        p[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}
}  // namespace

FGP_EXPORT int fgp_dbg_gemm_bench(int device, int M, int N, int K, int lower, int beta_one, int reps, double* ms_out,
                                  double* flops_out) {
    if (M % GEMM_BM || N % GEMM_BN || K % GEMM_KC || reps < 1 || !ms_out) return FGP_ERR_BAD_ARG;
    DeviceGuard dg(device);
    if (gemm_nt_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    double *dC = nullptr, *dA = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = FGP_OK;
    if (cudaMalloc(&dC, (size_t)M * N * 8) != cudaSuccess || cudaMalloc(&dA, (size_t)M * K * 8) != cudaSuccess ||
        cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        // random operands in [-1, 1): all-zero matrices would not toggle the datapath like real panels do
        fill_random_kernel<<<1024, 256>>>(dC, (int64_t)M * N, 0x9E3779B97F4A7C15ull);
        fill_random_kernel<<<1024, 256>>>(dA, (int64_t)M * K, 0xD1B54A32D192ED03ull);
        GemmArgs g{};
        g.C = dC; g.ldc = M;
        g.A = dA; g.lda = M;
        g.B = dA; g.ldb = M;  // SYRK-shaped: B = the first N rows of A
        g.M = M; g.N = N; g.K = K;
        g.alpha = -1.0; g.beta_one = beta_one; g.lower = lower; g.k_from_tile = 0;
        gemm_nt_launch(g, LaunchCtx{});  // warm-up
        cudaEventRecord(e0, nullptr);
        for (int r = 0; r < reps; ++r) gemm_nt_launch(g, LaunchCtx{});
        cudaEventRecord(e1, nullptr);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = FGP_ERR_CUDA;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / reps;
        if (flops_out) *flops_out = gemm_nt_flops(g);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(dC);
    cudaFree(dA);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    return rc;
}

FGP_EXPORT int fgp_dbg_gemm_nt(int device, double* C, int64_t ldc, const double* A, int64_t lda, const double* B,
                               int64_t ldb, int M, int N, int K, double alpha, int beta_one, int lower) {
    if (M % GEMM_BM || N % GEMM_BN || K % GEMM_KC) return FGP_ERR_BAD_ARG;
    DeviceGuard dg(device);
    if (gemm_nt_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    double *dC = nullptr, *dA = nullptr, *dB = nullptr;
    int rc = FGP_OK;
    if (cudaMalloc(&dC, (size_t)M * N * 8) != cudaSuccess || cudaMalloc(&dA, (size_t)M * K * 8) != cudaSuccess ||
        cudaMalloc(&dB, (size_t)N * K * 8) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        cudaMemcpy2D(dC, (size_t)M * 8, C, (size_t)ldc * 8, (size_t)M * 8, N, cudaMemcpyHostToDevice);
        cudaMemcpy2D(dA, (size_t)M * 8, A, (size_t)lda * 8, (size_t)M * 8, K, cudaMemcpyHostToDevice);
        cudaMemcpy2D(dB, (size_t)N * 8, B, (size_t)ldb * 8, (size_t)N * 8, K, cudaMemcpyHostToDevice);
        GemmArgs g{};
        g.C = dC; g.ldc = M;
        g.A = dA; g.lda = M;
        g.B = dB; g.ldb = N;
        g.M = M; g.N = N; g.K = K;
        g.alpha = alpha; g.beta_one = beta_one; g.lower = lower; g.k_from_tile = 0;
        gemm_nt_launch(g, LaunchCtx{});
        if (cudaDeviceSynchronize() != cudaSuccess) rc = FGP_ERR_CUDA;
        cudaMemcpy2D(C, (size_t)ldc * 8, dC, (size_t)M * 8, (size_t)M * 8, N, cudaMemcpyDeviceToHost);
    }
    cudaFree(dC);
    cudaFree(dA);
    cudaFree(dB);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    return rc;
}

// test hook: C(lower or full, M x M) -= A A^T through the tcgen05 exact-integer path (csrc/ozaki.cu) on device copies of
// host matrices; A is M x K.  lbo / sbo <= 0: the production descriptor offsets.
FGP_EXPORT int fgp_dbg_ozaki_syrk(int device, double* C, int64_t ldc, const double* A, int64_t lda, int M, int K, int lower,
                                  int row_skip, int tiles_per_cta, int lbo, int sbo) {
    if (M % 128 || K % 128 || M <= 0 || K <= 0 || K > 512) return FGP_ERR_BAD_ARG;
    DeviceGuard dg(device);
    if (ozaki_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    double *dC = nullptr, *dA = nullptr, *dS = nullptr;
    int8_t* dD = nullptr;
    int rc = FGP_OK;
    if (cudaMalloc(&dC, (size_t)M * M * 8) != cudaSuccess || cudaMalloc(&dA, (size_t)M * K * 8) != cudaSuccess ||
        cudaMalloc(&dS, (size_t)M * 8) != cudaSuccess || cudaMalloc(&dD, ozaki_slice_bytes(M, K)) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        cudaMemcpy2D(dC, (size_t)M * 8, C, (size_t)ldc * 8, (size_t)M * 8, M, cudaMemcpyHostToDevice);
        cudaMemcpy2D(dA, (size_t)M * 8, A, (size_t)lda * 8, (size_t)M * 8, K, cudaMemcpyHostToDevice);
        GemmArgs g{};
        g.C = dC; g.ldc = M;
        g.M = M; g.N = M; g.K = K;
        g.alpha = -1.0; g.beta_one = 1; g.lower = lower; g.row_skip = row_skip;
        ozaki_slice_launch(dA, M, M, K, dD, dS, LaunchCtx{});
        ozaki_update_launch(g, dD, dS, dD, dS, tiles_per_cta, LaunchCtx{}, lbo > 0 ? (uint32_t)lbo : OZ_LBO, sbo > 0 ? (uint32_t)sbo : OZ_SBO);
        if (cudaDeviceSynchronize() != cudaSuccess) rc = FGP_ERR_CUDA;
        cudaMemcpy2D(C, (size_t)ldc * 8, dC, (size_t)M * 8, (size_t)M * 8, M, cudaMemcpyDeviceToHost);
    }
    cudaFree(dC);
    cudaFree(dA);
    cudaFree(dS);
    cudaFree(dD);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    if (gemm_nt_take_error()) rc = FGP_ERR_CUDA;
    return rc;
}

// measurement hook: `reps` launches of the slicing kernel and of the tcgen05 update C(lower, M x M) -= A A^T on device-resident
// random data (A is M x K); CUDA-event time per launch of each
FGP_EXPORT void fgp_dbg_ozaki_experiment(int flags) { ozaki_set_experiment(flags); }

FGP_EXPORT int fgp_dbg_ozaki_bench(int device, int M, int K, int reps, int tiles_per_cta, double* ms_update, double* ms_slice) {
    if (M % 128 || K % 128 || M <= 0 || K <= 0 || K > 512 || reps < 1 || !ms_update) return FGP_ERR_BAD_ARG;
    DeviceGuard dg(device);
    if (ozaki_prepare() != cudaSuccess) return FGP_ERR_CUDA;
    double *dC = nullptr, *dA = nullptr, *dS = nullptr;
    int8_t* dD = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    int rc = FGP_OK;
    if (cudaMalloc(&dC, (size_t)M * M * 8) != cudaSuccess || cudaMalloc(&dA, (size_t)M * K * 8) != cudaSuccess ||
        cudaMalloc(&dS, (size_t)M * 8) != cudaSuccess || cudaMalloc(&dD, ozaki_slice_bytes(M, K)) != cudaSuccess ||
        cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess || cudaEventCreate(&e2) != cudaSuccess)
        rc = FGP_ERR_CUDA;
    if (rc == FGP_OK) {
        fill_random_kernel<<<1024, 256>>>(dC, (int64_t)M * M, 0x9E3779B97F4A7C15ull);
        fill_random_kernel<<<1024, 256>>>(dA, (int64_t)M * K, 0xD1B54A32D192ED03ull);
        GemmArgs g{};
        g.C = dC; g.ldc = M;
        g.M = M; g.N = M; g.K = K;
        g.alpha = -1.0; g.beta_one = 1; g.lower = 1;
        ozaki_slice_launch(dA, M, M, K, dD, dS, LaunchCtx{});
        ozaki_update_launch(g, dD, dS, dD, dS, tiles_per_cta, LaunchCtx{});
        cudaEventRecord(e0, nullptr);
        for (int r = 0; r < reps; ++r) ozaki_slice_launch(dA, M, M, K, dD, dS, LaunchCtx{});
        cudaEventRecord(e1, nullptr);
        for (int r = 0; r < reps; ++r) ozaki_update_launch(g, dD, dS, dD, dS, tiles_per_cta, LaunchCtx{});
        cudaEventRecord(e2, nullptr);
        if (cudaEventSynchronize(e2) != cudaSuccess) rc = FGP_ERR_CUDA;
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        if (ms_slice) *ms_slice = a / reps;
        *ms_update = b / reps;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    cudaFree(dC);
    cudaFree(dA);
    cudaFree(dS);
    cudaFree(dD);
    if (cudaGetLastError() != cudaSuccess) rc = FGP_ERR_CUDA;
    if (gemm_nt_take_error()) rc = FGP_ERR_CUDA;
    return rc;
}
