// Dependent-issue latencies of the scalar fp64 path on sm_100a (one warp, clock64 around N dependent ops).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void lat_kernel(double* out, long long* clk, double seed) {
    __shared__ double sm[64];
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 0.0;
    sm[threadIdx.x] = x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) x = fma(x, y, 1e-9);                                   // DFMA chain
            if (MODE == 1) x = x * y;                                             // DMUL chain
            if (MODE == 2) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);  // 64-bit shuffle (2 SHFL) chain
            if (MODE == 3) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }  // MUFU.RSQ64H chain
            if (MODE == 4) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31]; }       // STS -> LDS round trip
            if (MODE == 5) dmma884(x, z, y, y);                                   // DMMA chain (same accumulator)
            if (MODE == 6) x = (x > 1e-300) ? x : y;                              // DSETP + select
            if (MODE == 7) x = x + y;                                             // DADD chain
            if (MODE == 8) x = sm[((int)__double2loint(x) & 31)];                 // dependent LDS
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x + z;
    if (threadIdx.x == 0) clk[MODE] = t1 - t0;
}

int main() {
    double* out;
    long long* clk;
    cudaMalloc(&out, 64 * 8);
    cudaMalloc(&clk, 16 * 8);
    const char* names[] = {"DFMA", "DMUL", "SHFL64", "MUFU.RSQ64H", "STS+LDS", "DMMA", "DSETP+SEL", "DADD", "LDS(dep)"};
    for (int rep = 0; rep < 2; ++rep) {
        lat_kernel<0><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<1><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<2><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<3><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<4><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<5><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<6><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<7><<<1, 32>>>(out, clk, 1.0);
        lat_kernel<8><<<1, 32>>>(out, clk, 1.0);
        cudaDeviceSynchronize();
    }
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"dependent_latency_cycles\":{");
    for (int m = 0; m < 9; ++m) printf("%s\"%s\":%.1f", m ? "," : "", names[m], h[m] / 1024.0);
    printf("}}\nerr: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
