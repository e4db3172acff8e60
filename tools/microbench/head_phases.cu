// Phase timing of potrf_head_kernel's diagonal-tile CTA: clock64 stamps of lane 0 of each of its 8 warps after every phase.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DFGP_HEAD_TIMING -o head_phases head_phases.cu
#include "../../friedrich_b200/csrc/potrf_head.cu"

#include <cmath>
#include <cstdio>
#include <vector>

int main(int argc, char** argv) {
    using namespace fgp;
    const int nt = argc > 1 ? atoi(argv[1]) : 1;
    if (potrf_head_prepare() != cudaSuccess) { printf("prepare failed\n"); return 1; }
    const int n = 128 * nt;
    std::vector<double> A((size_t)n * n);
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) A[r + (size_t)c * n] = (r == c ? 2.0 : 0.0) + exp(-0.01 * (r - c) * (r - c));
    double *dA, *dinv, *dW, *dP;
    int *info, *sync;
    cudaMalloc(&dA, (size_t)n * n * 8); cudaMalloc(&dinv, (size_t)n * 128 * 8); cudaMalloc(&dW, 512 * 512 * 8);
    cudaMalloc(&dP, 512 * 512 * 8); cudaMalloc(&info, 4); cudaMalloc(&sync, HEAD_SYNC_INTS * 4);
    cudaMemset(info, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice);
        cudaMemset(sync, 0, HEAD_SYNC_INTS * 4);
        cudaEventRecord(e0);
        launch_potrf_head(dA, n, nt, dinv, dW, dP, sync, 0, 0.0, info, 0, LaunchCtx{});
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep < 2) continue;
        long long clk[8 * 64];
        cudaMemcpyFromSymbol(clk, g_head_clk, sizeof(clk));
        const long long t0 = clk[0];  // warp 0, slot 0: start of the LAST diagonal tile (each tile overwrites the slots)
        printf("{\"nt\":%d,\"us\":%.1f,\"last_tile\":\"cycles since tile start, per warp (w0 = pivot warp)\"}\n", nt, ms * 1e3);
        const char* names[32] = {"start", "loaded"};
        static char buf[32][24];
        for (int s = 0; s < 4; ++s) {
            snprintf(buf[2 + 6 * s], 24, "piv%d", s); snprintf(buf[3 + 6 * s], 24, "x32/inv%d", s);
            snprintf(buf[4 + 6 * s], 24, "A%d start", s); snprintf(buf[5 + 6 * s], 24, "B%d start", s);
            snprintf(buf[6 + 6 * s], 24, "C%d start", s); snprintf(buf[7 + 6 * s], 24, "grp%d done", s);
            for (int k = 2; k < 8; ++k) names[k + 6 * s] = buf[k + 6 * s];
        }
        names[26] = "loop_end"; names[27] = "apply3"; names[28] = "stored";
        for (int slot = 0; slot < 29; ++slot) {
            if (!names[slot]) continue;
            printf("  %-10s", names[slot]);
            for (int w = 0; w < 8; ++w) {
                const long long v = clk[64 * w + slot];
                if (v >= t0 && v - t0 < 10000000) printf(" %7lld", v - t0); else printf(" %7s", "-");
            }
            printf("\n");
        }
    }
    {
        long long clk[8 * 64];
        cudaMemcpyFromSymbol(clk, g_head_clk, sizeof(clk));
        const long long t0 = clk[0];
        // warp 0 inside the LAST sub-panel (s = 3): pivot32 panels (chain end / rank-8 update end) and head_x32 stages
        printf("  last pivot32 (s=3), since tile start: ");
        for (int p = 0; p < 4; ++p) printf("chain%d %lld upd%d %lld  ", p, clk[34 + 2 * p] - t0, p, clk[35 + 2 * p] - t0);
        printf("\n  last head_x32: inv8 done %lld, 8->16 done %lld\n", clk[32] - t0, clk[33] - t0);
    }
    int h_info = 0;
    cudaMemcpy(&h_info, info, 4, cudaMemcpyDeviceToHost);
    printf("info=%d err: %s\n", h_info, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
