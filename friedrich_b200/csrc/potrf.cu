// potrf.cu — diagonal-tile factorisation kernel and the blocked lower Cholesky driver (contract in potrf.cuh).
#include "potrf.cuh"

namespace fgp {

// One CTA: factor the 128x128 diagonal tile in shared memory (4 sub-panels of 32 columns: the 32x32 pivot block
// is factored by one warp with the pivot row broadcast by shuffle), write L back, then invert L in place
// (blocked dtrtri) and write the inverse (upper part zeroed) to `inv` (ld = 128).
__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ inv, double* __restrict__ invT, int has_sub,
                  double sub, int* info, int col_base) {
    extern __shared__ __align__(16) double dsm[];
    double* T = dsm;                       // [128][DIAG_DS], row-major: T[r*DS + c]
    double* W = dsm + 128 * DIAG_DS;       // [96][33] scratch
    double* rdiag = W + 96 * 33;           // 1 / L[k][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < 128 * 128; idx += 256) {
        const int r = idx & 127, c = idx >> 7;
        T[r * DIAG_DS + c] = (r >= c) ? A[r + (int64_t)c * lda] : 0.0;
    }
    __syncthreads();

    for (int s = 0; s < 4; ++s) {
        const int c0 = 32 * s;
        // (a) pivot block: lane r owns row c0 + r
        if (warp == 0) {
            double row[32];
            double* Tr = T + (c0 + lane) * DIAG_DS + c0;
#pragma unroll
            for (int k = 0; k < 32; ++k) row[k] = Tr[k];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                double dk = __shfl_sync(0xffffffffu, row[k], k);
                if (!(dk > 0.0)) {  // zero, negative or NaN pivot (nalgebra: is_zero / try_sqrt fails)
                    if (has_sub && sub > 0.0) dk = sub;
                    else {
                        if (lane == 0) atomicCAS(info, 0, col_base + c0 + k + 1);
                        dk = nan("");
                    }
                }
                // 1/sqrt first (one MUFU seed + Newton steps): the column scaling, which the next pivot waits for, needs
                // only rs; sqrt(dk) = dk * rs gets one correction step off the critical path (<= 1 ulp)
                const double rs = rsqrt(dk);
                double sk = dk * rs;
                sk = fma(0.5 * rs, fma(-sk, sk, dk), sk);
                if (lane > k) row[k] *= rs;
                else if (lane == k) { row[k] = sk; rdiag[c0 + k] = rs; }
                if (lane >= k) Tr[k] = row[k];
                __syncwarp();
#pragma unroll
                for (int c = k + 1; c < 32; ++c) {
                    const double lck = T[(c0 + c) * DIAG_DS + c0 + k];
                    row[c] = fma(-row[k], lck, row[c]);
                }
            }
        }
        __syncthreads();
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
        // (c) trailing update inside the tile: A[r][c] -= sum_k X[r][k] X[c][k]
        for (int e = tid; e < tn * tn; e += 256) {
            const int rr = e / tn, cc = e - rr * tn;
            if (cc <= rr) {
                const double* xr = T + (c0 + 32 + rr) * DIAG_DS + c0;
                const double* xc = T + (c0 + 32 + cc) * DIAG_DS + c0;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    a0 = fma(xr[k], xc[k], a0);
                    a1 = fma(xr[k + 1], xc[k + 1], a1);
                    a2 = fma(xr[k + 2], xc[k + 2], a2);
                    a3 = fma(xr[k + 3], xc[k + 3], a3);
                }
                T[(c0 + 32 + rr) * DIAG_DS + c0 + 32 + cc] -= (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();
    }

    // the factor goes back to the matrix (lower part only)
    for (int idx = tid; idx < 128 * 128; idx += 256) {
        const int r = idx & 127, c = idx >> 7;
        if (r >= c) A[r + (int64_t)c * lda] = T[r * DIAG_DS + c];
    }
    __syncthreads();

    // in-place inverse of the lower-triangular tile, 32-blocks from the last to the first (LAPACK dtrtri, lower)
    for (int s = 3; s >= 0; --s) {
        const int c0 = 32 * s;
        const int tn = 128 - c0 - 32;
        // (1) invert the 32x32 diagonal block: lane c owns column c of the inverse
        if (warp == 0) {
            double x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                x[i] *= rdiag[c0 + i];
#pragma unroll
                for (int i2 = i + 1; i2 < 32; ++i2) x[i2] = fma(-x[i], T[(c0 + i2) * DIAG_DS + c0 + i], x[i2]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i >= lane) T[(c0 + i) * DIAG_DS + c0 + lane] = x[i];
        }
        __syncthreads();
        if (tn > 0) {
            // (2a) W = Ainv_trailing * C  (C = original block below the pivot block)
            for (int e = tid; e < tn * 32; e += 256) {
                const int rr = e >> 5, k = e & 31;
                const double* ar = T + (c0 + 32 + rr) * DIAG_DS + c0 + 32;
                double a0 = 0, a1 = 0;
                int m = 0;
                for (; m + 1 <= rr; m += 2) {
                    a0 = fma(ar[m], T[(c0 + 32 + m) * DIAG_DS + c0 + k], a0);
                    a1 = fma(ar[m + 1], T[(c0 + 32 + m + 1) * DIAG_DS + c0 + k], a1);
                }
                if (m <= rr) a0 = fma(ar[m], T[(c0 + 32 + m) * DIAG_DS + c0 + k], a0);
                W[rr * 33 + k] = a0 + a1;
            }
            __syncthreads();
            // (2b) C = -W * Dinv
            for (int e = tid; e < tn * 32; e += 256) {
                const int rr = e >> 5, k = e & 31;
                double a0 = 0;
                for (int m = k; m < 32; ++m) a0 = fma(W[rr * 33 + m], T[(c0 + m) * DIAG_DS + c0 + k], a0);
                T[(c0 + 32 + rr) * DIAG_DS + c0 + k] = -a0;
            }
            __syncthreads();
        }
    }
    for (int idx = tid; idx < 128 * 128; idx += 256) {
        const int r = idx & 127, c = idx >> 7;
        inv[r + c * 128] = (r >= c) ? T[r * DIAG_DS + c] : 0.0;
    }
    for (int idx = tid; idx < 128 * 128; idx += 256) {  // transposed copy for the adjoint solves
        const int c = idx & 127, r = idx >> 7;
        invT[c + r * 128] = (r >= c) ? T[r * DIAG_DS + c] : 0.0;
    }
}

cudaError_t potrf_prepare() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = gemm_nt_prepare();
    if (e == cudaSuccess) done = true;
    return e;
}

// block columns [J, Jend) of one panel, left-looking inside the panel; every launch goes to c.st
void factor_panel(double* A, int64_t lda, int64_t np, int64_t J, int64_t Jend, double* invdiag, double* invdiagT,
                         int has_sub, double sub, int* info, const LaunchCtx& c, PotrfCounters* cnt) {
    const int64_t nb = np / TILE;
    for (int64_t j = J; j < Jend; ++j) {
        double* Ajj = A + j * TILE + j * TILE * lda;
        if (j > J) {
            GemmArgs g{};
            g.C = Ajj; g.ldc = lda;
            g.A = A + j * TILE + J * TILE * lda; g.lda = lda;
            g.B = g.A; g.ldb = lda;
            g.M = (int)(np - j * TILE); g.N = TILE; g.K = (int)((j - J) * TILE);
            g.alpha = -1.0; g.beta_one = 1; g.lower = 0; g.k_from_tile = 0;
            cnt->launches += gemm_nt_launch(g, c) > 0;
        }
        {
            ProfScope ps(c, PROF_POTRF_DIAG, 128.0 * 128.0 * 128.0 / 3.0);
            potrf_diag_kernel<<<1, 256, DIAG_SMEM_BYTES, c.st>>>(Ajj, lda, invdiag + j * TILE * TILE,
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

void potrf_lower(double* A, int64_t lda, int64_t np, int64_t jb_begin, double* invdiag, double* invdiagT, int has_sub,
                 double sub, int* info, const LaunchCtx& st, const PotrfLookahead* la, PotrfCounters* cnt) {
    const int64_t nb = np / TILE;
    if (jb_begin >= nb) return;
    if (!la || !la->panel) {
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
    factor_panel(A, lda, np, J, Jend, invdiag, invdiagT, has_sub, sub, info, pc, cnt);
    cudaEventRecord(la->ev_panel, la->panel);
    bool first = true;
    while (Jend < nb) {
        const int64_t Jend2 = (Jend + PANEL_TILES < nb) ? Jend + PANEL_TILES : nb;
        cudaStreamWaitEvent(st.st, la->ev_panel, 0);                   // M: panel [J, Jend) is final
        if (!first) cudaStreamWaitEvent(la->panel, la->ev_trail, 0);   // P: previous trailing update reached columns >= Jend
        trailing_update(A, lda, np, J, Jend, Jend, Jend2, pc, cnt);    // look-ahead columns
        factor_panel(A, lda, np, Jend, Jend2, invdiag, invdiagT, has_sub, sub, info, pc, cnt);
        cudaEventRecord(la->ev_panel, la->panel);
        trailing_update(A, lda, np, J, Jend, Jend2, nb, st, cnt);      // the rest
        cudaEventRecord(la->ev_trail, st.st);
        first = false;
        J = Jend;
        Jend = Jend2;
    }
    cudaStreamWaitEvent(st.st, la->ev_panel, 0);  // join: everything after the factorisation runs on M
}

}  // namespace fgp
