// model.cuh — the device-resident state behind an `fgp_model*` handle and the small host helpers every entry point
// shares.  Mirrors the private fields of the reference's `GaussianProcess` (src/gaussian_process/mod.rs:58-79):
// training_inputs (EMatrix, capacity-padded), training_outputs (EVector, already minus the prior) and
// covmat_cholesky — plus what the device path caches on top (alpha = K^-1 y, z = L^-1 y, inverted diagonal blocks).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fgp.h"
#include "common.cuh"
#include "gemm_nt.cuh"

namespace fgp {

struct DevBuf {
    double* p = nullptr;
    size_t cap = 0;  // in doubles
    // grow-only; `keep` preserves the old contents (first `cap` doubles); `zero` clears whatever is new (buffers whose
    // never-written parts are read as zeros: the strict upper triangles of the inverse diagonal tiles)
    cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0, bool zero = false) {
        if (n <= cap) return cudaSuccess;
        double* q = nullptr;
        cudaError_t e = cudaMalloc(&q, n * sizeof(double));
        if (e != cudaSuccess) return e;
        const size_t kept = (keep && p) ? cap : 0;
        if (zero) e = cudaMemsetAsync(q + kept, 0, (n - kept) * sizeof(double), st);
        if (e == cudaSuccess && kept) e = cudaMemcpyAsync(q, p, kept * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && (kept || zero)) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(q);
            return e;
        }
        if (p) cudaFree(p);
        p = q;
        cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace fgp

struct fgp_comm;  // multi-GPU transport (comm.cuh)

struct fgp_model {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;  // look-ahead / copy stream
    cudaStream_t st3 = nullptr;  // side stream of the look-ahead: the next panel's block columns 1.. are updated here while
                                 // block column 0 is already being factored on st2
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr, evC = nullptr, evD = nullptr;
    std::mutex mu;
    std::string err;

    // training state ------------------------------------------------------------------------------------------
    bool fitted = false;
    int64_t n = 0, d = 0, dp = 0, np = 0;  // samples, dims, padded dims (multiple of 4), padded samples (multiple of 128)
    int64_t cap = 0;                       // capacity in rows (multiple of 128) = leading dimension of L
    fgp::DevBuf xr, xc, nc, nr, cmean;     // raw / centred points [cap][dp], squared norms, column means [dp]
    fgp::DevBuf y, z, alpha, work;         // outputs (0 padded), L^-1 y, K^-1 y, scratch vector
    fgp::DevBuf L, inv, invT;              // factor [cap x cap], inverse diagonal blocks and their transposes
    // head schedule of the blocked Cholesky (potrf.cuh PotrfWork): per-panel inverse blocks W, head scratch, panel buffers
    fgp::DevBuf Wp, Pscr, pbuf[2];
    fgp::DevBuf ozDigits, ozScale;         // base-128 digit slices + row scales of the panel being applied (csrc/ozaki.cuh)
    bool tcgen05 = true;                   // FGP_OPT_TCGEN05
    int shard_pipe = -1;                   // FGP_OPT_SHARD_PIPE: row-piece schedule of the multi-GPU fit (-1 = automatic: from 3 ranks up)
    fgp::DevBuf ozL, ozLscale;             // digit slices / row scales of EVERY panel of L with >= OZ_MIN_ROWS rows below it, kept by a
    std::vector<int64_t> ozOffBytes, ozOffRows;   // full single-GPU fit for the solves of predict: per panel byte / row offset, -1 = none
    bool ozL_valid = false;
    fgp::DevBuf ozU;                       // digit slices of U = L^-T for K^-1 = U U^T on tcgen05 (LML gradient)
    int* head_sync = nullptr;              // [head_sync_cap][HEAD_SYNC_INTS]
    int64_t head_sync_cap = 0;
    std::vector<int64_t> pstart;           // first block column of every panel of the current factor (W slot = index)
    bool w_valid = false;                  // Wp holds the inverse diagonal block of EVERY panel in pstart (single-GPU head fits)
    cudaEvent_t evTop = nullptr, evRest = nullptr, evCopy[2] = {nullptr, nullptr};
    bool head_schedule = true;             // FGP_OPT_HEAD: 0 = the per-block-column schedule of round 1 (A/B runs)
    fgp::DevBuf staging;                   // H2D landing zone (column-major inputs)
    int* info_d = nullptr;
    int* info_h = nullptr;                 // pinned
    int64_t failed_col = -1;

    // query state ---------------------------------------------------------------------------------------------
    int64_t q = 0, qp = 0;
    fgp::DevBuf qr, qc, qnc, qnr;          // query points and norms
    fgp::DevBuf bt;                        // transposed cross-covariance / solve buffer [qp x np]
    fgp::DevBuf partial, mean_d, var_d, scalars, kqq;
    double* pinned = nullptr;              // host staging for results
    size_t pinned_cap = 0;
    bool have_mean = false, have_var = false;
    bool force_batched_predict = false;    // fgp_predict_cov needs the transposed solve buffer whatever q is

    // LML workspace -------------------------------------------------------------------------------------------
    fgp::DevBuf U, Kinv, lml_partial, lml_rows;
    bool kinv_valid = false;               // Kinv holds K^-1 of the CURRENT factor (set by fgp_lml_gradient, cleared by every refit)

    // multi-GPU -----------------------------------------------------------------------------------------------
    fgp_comm* comm = nullptr;

    // measurement ---------------------------------------------------------------------------------------------
    float last_ms = 0.f;
    int64_t launches = 0;
    bool profiling = false;
    bool lookahead = true;                 // two-stream panel look-ahead in the blocked Cholesky (fgp_set_option)
    fgp::Profiler prof;
    fgp::LaunchCtx ctx() { return fgp::LaunchCtx{st, profiling ? &prof : nullptr}; }
};

namespace fgp {

inline int fail(fgp_model* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    return code;
}

#define CU(m, call)                                                                                                 \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess)                                                                                     \
            return fgp::fail(m, FGP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__   \
                                                  ":" + std::to_string(__LINE__) + ")");                             \
    } while (0)

#define FGP_TRY(expr)                   \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != FGP_OK) return rc__; \
    } while (0)

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

inline int ensure_pinned(fgp_model* m, size_t doubles) {
    if (doubles <= m->pinned_cap) return FGP_OK;
    if (m->pinned) cudaFreeHost(m->pinned);
    m->pinned = nullptr;
    m->pinned_cap = 0;
    CU(m, cudaMallocHost(&m->pinned, doubles * sizeof(double)));
    m->pinned_cap = doubles;
    return FGP_OK;
}

inline void begin_timed(fgp_model* m) {
    m->launches = 0;
    if (m->profiling) m->prof.reset();
    cudaEventRecord(m->ev0, m->st);
}
inline int end_timed(fgp_model* m) {
    CU(m, cudaEventRecord(m->ev1, m->st));
    CU(m, cudaEventSynchronize(m->ev1));
    CU(m, cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));
    if (m->profiling) m->prof.collect();
    CU(m, cudaGetLastError());
    if (gemm_nt_take_error()) return fgp::fail(m, FGP_ERR_CUDA, "kernel launch setup failed (cuTensorMapEncodeTiled)");
    return FGP_OK;
}

}  // namespace fgp
