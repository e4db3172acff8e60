// Phase timing of potrf_diag_kernel (clock64 at phase boundaries, thread 0).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFGP_DIAG_TIMING -o diag_phases diag_phases.cu
#include "../../friedrich_b200/csrc/gemm_nt.cu"
#include "../../friedrich_b200/csrc/potrf.cu"

#include <cstdio>
#include <vector>

int main() {
    using namespace fgp;
    if (potrf_prepare() != cudaSuccess) { printf("prepare failed\n"); return 1; }
    const int n = 128;
    std::vector<double> A(n * n);
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) A[r + c * n] = (r == c ? 2.0 : 0.0) + exp(-0.01 * (r - c) * (r - c));
    double *dA, *dinv, *dinvT;
    int* info;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dinv, n * n * 8); cudaMalloc(&dinvT, n * n * 8); cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM_BYTES>>>(dA, n, dinv, dinvT, 0, 0.0, info, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long clk[32];
        cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof(clk));
        printf("{\"rep\":%d,\"us\":%.1f,\"cycles\":{\"load\":%lld", rep, ms * 1e3, clk[1] - clk[0]);
        for (int s = 0; s < 4; ++s)
            printf(",\"pivot%d\":%lld,\"below%d\":%lld,\"syrk%d\":%lld", s, clk[2 + 3 * s] - clk[1 + 3 * s], s,
                   clk[3 + 3 * s] - clk[2 + 3 * s], s, clk[4 + 3 * s] - clk[3 + 3 * s]);
        printf(",\"store\":%lld,\"inv_rowblock3\":%lld,\"store_inv\":%lld,\"total\":%lld}}\n",
               clk[14] - clk[13], clk[18] - clk[14], clk[19] - clk[18], clk[19] - clk[0]);
        // verify: L L^T = A (lower) and inv * L = I
        std::vector<double> L(n * n), X(n * n), XT(n * n);
        cudaMemcpy(L.data(), dA, n * n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(X.data(), dinv, n * n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(XT.data(), dinvT, n * n * 8, cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, e3 = 0;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) {
                double s1 = 0, s2 = 0;
                for (int k = 0; k <= c; ++k) s1 += L[r + k * n] * L[c + k * n];
                for (int k = c; k <= r; ++k) s2 += X[r + k * n] * L[k + c * n];
                e1 = fmax(e1, fabs(s1 - A[r + c * n]));
                e2 = fmax(e2, fabs(s2 - (r == c ? 1.0 : 0.0)));
                e3 = fmax(e3, fabs(X[r + c * n] - XT[c + r * n]));
            }
        printf("{\"check\":{\"max|LLt-A|\":%.3e,\"max|XL-I|\":%.3e,\"max|X-XT^T|\":%.3e}}\n", e1, e2, e3);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
