// fp64 peak microbenchmarks for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) register-resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Output: one JSON line per test. Used to set the fp64 roofline denominator (no fp64 figure in MEASURED_PEAKS.json).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NACC independent accumulator tiles per warp; MODE 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
template <int MODE, int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double av, double bv) {
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = av + i * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = bv + i * 1e-3;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = 0; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) dmma884(c[i][0], c[i][1], a[0], b[0]);
            if (MODE == 1) dmma1684(c[i], a, b[0]);
            if (MODE == 2) dmma1688(c[i], a, b);
            if (MODE == 3) dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"l2_bytes\":%d,\"smem_optin\":%zu,\"regs_per_sm\":%d,\"cc\":\"%d.%d\"}\n",
           p.name, p.multiProcessorCount, clk, p.l2CacheSize, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor, p.major, p.minor);
    double* out; CK(cudaMalloc(&out, 8));
    int sms = p.multiProcessorCount;
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            {
                float ms = time_ms([&] { dfma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
                double fl = 2.0 * 8 * iters * (double)threads * grid;
                printf("{\"test\":\"dfma\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, bps, ms, fl / ms * 1e-9);
            }
#define RUN_DMMA(MODE, NACC, MK, name) { \
                float ms = time_ms([&] { dmma_kernel<MODE, NACC><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5); \
                double fl = 2.0 * (MK) * NACC * iters * (double)(threads / 32) * grid; \
                printf("{\"test\":\"%s\",\"nacc\":%d,\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", name, NACC, threads, bps, ms, fl / ms * 1e-9); }
            RUN_DMMA(0, 4, 8 * 8 * 4, "dmma_m8n8k4");
            RUN_DMMA(0, 16, 8 * 8 * 4, "dmma_m8n8k4");
            RUN_DMMA(1, 8, 16 * 8 * 4, "dmma_m16n8k4");
            RUN_DMMA(2, 8, 16 * 8 * 8, "dmma_m16n8k8");
            RUN_DMMA(3, 8, 16 * 8 * 16, "dmma_m16n8k16");
        }
    }
    // sustained run (3 s) of the best shape to see clocks under load
    {
        int grid = sms * 2, threads = 512;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0; float ms = 0;
        while (ms < 3000.f) {
            for (int i = 0; i < 10; ++i) dmma_kernel<0, 16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
            launches += 10;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        double fl = 2.0 * 256 * 16 * iters * (double)(threads / 32) * grid * launches;
        printf("{\"test\":\"dmma_m8n8k4_sustained\",\"ms\":%.1f,\"tflops\":%.3f}\n", ms, fl / ms * 1e-9);
        CK(cudaEventRecord(e0)); launches = 0; ms = 0;
        while (ms < 3000.f) {
            for (int i = 0; i < 10; ++i) dfma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
            launches += 10;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        fl = 2.0 * 8 * iters * (double)threads * grid * launches;
        printf("{\"test\":\"dfma_sustained\",\"ms\":%.1f,\"tflops\":%.3f}\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
