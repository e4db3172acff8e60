// Where does the production GEMM kernel lose time?  Compiles friedrich_b200/csrc/gemm_nt.cu with -DFGP_GEMM_EXP=n and times
// SYRK-shaped launches.  Build (one binary per variant):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFGP_GEMM_EXP=n -o gemm_variant_n gemm_variants.cu
#include "../../friedrich_b200/csrc/gemm_nt.cu"

#include <cstdio>

int main(int argc, char** argv) {
    using namespace fgp;
    if (gemm_nt_prepare() != cudaSuccess) { printf("prepare failed\n"); return 1; }
    const int shapes[][3] = {{16384, 16384, 512}, {16384, 16384, 1024}, {16384, 16384, 128}, {4096, 4096, 512}};
    for (auto& sh : shapes) {
        const int M = sh[0], N = sh[1], K = sh[2];
        double *C, *A;
        cudaMalloc(&C, (size_t)M * N * 8);
        cudaMalloc(&A, (size_t)M * K * 8);
        cudaMemset(C, 0, (size_t)M * N * 8);
        cudaMemset(A, 0, (size_t)M * K * 8);
        GemmArgs g{};
        g.C = C; g.ldc = M; g.A = A; g.lda = M; g.B = A; g.ldb = M;
        g.M = M; g.N = N; g.K = K; g.alpha = -1.0; g.beta_one = 1; g.lower = 1;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        gemm_nt_launch(g, LaunchCtx{});
        cudaEventRecord(e0);
        const int reps = 5;
        for (int r = 0; r < reps; ++r) gemm_nt_launch(g, LaunchCtx{});
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("{\"exp\":%d,\"M\":%d,\"K\":%d,\"ms\":%.4f,\"tflops\":%.3f,\"err\":\"%s\"}\n", FGP_GEMM_EXP, M, K, ms,
               gemm_nt_flops(g) / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
        cudaFree(C); cudaFree(A);
    }
    return 0;
}
