// fp64 pipe study for B200 (sm_100a): what bounds a DMMA.8x8x4 main loop, and is DFMA a separate pipe?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
// Output: one JSON line per test (profiles/fp64_pipes_r01.jsonl).  Tests:
//   dmma_same      NACC independent accumulators, one shared (a, b) operand pair (the r01 peak test), long and short runs
//   dmma_tile      GEMM-like register pattern: MI x NI accumulators, MI a-operands, NI b-operands (register resident)
//   dmma_smem      as dmma_tile with the operands re-read from shared memory every k-step (our GEMM main loop w/o TMA)
//   dfma_tile      8x8 DFMA register tile, operands from shared memory (LDS.128): the ceiling of a vector-pipe GEMM
//   mix            half of the warps run dmma_same, half run DFMA chains: do the two add up?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, int TH>
__global__ void __launch_bounds__(TH) dmma_same(double* out, int iters, double av, double bv) {
    double c[NACC][2];
    double a = av + threadIdx.x * 1e-6, b = bv + threadIdx.x * 1e-7;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int MI, int NI, int TH>
__global__ void __launch_bounds__(TH) dmma_tile(double* out, int iters, double av, double bv) {
    double c[MI][NI][2];
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = av + i * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < NI; ++i) b[i] = bv + i * 1e-3;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { c[i][j][0] = i; c[i][j][1] = j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// operands from shared memory, [k][132] layout like gemm_nt (column stride 132 doubles); KC k-values per "stage", the
// warp re-reads the same stage every iteration (no global traffic)
template <int MI, int NI, int TH>
__global__ void __launch_bounds__(TH) dmma_smem(double* out, int iters, double av) {
    extern __shared__ double sm[];
    constexpr int LDS_ = 132, KC = 16;
    for (int i = threadIdx.x; i < 2 * KC * LDS_; i += blockDim.x) sm[i] = av + (i % 97) * 1e-4;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, t = lane & 3;
    const double* As = sm + t * LDS_ + (8 * MI * (warp & 1)) % (128 - 8 * MI + 1) + gq;
    const double* Bs = sm + KC * LDS_ + t * LDS_ + (8 * NI * (warp >> 1)) % (128 - 8 * NI + 1) + gq;
    double c[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { c[i][j][0] = i; c[i][j][1] = j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < KC / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = As[kk * 4 * LDS_ + 8 * i];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = Bs[kk * 4 * LDS_ + 8 * j];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

// 8x8 DFMA register tile; thread (tx, ty) of a 16x16 thread block tile; A, B panels [k][128] in smem; LDS.128 pairs.
__global__ void __launch_bounds__(256) dfma_tile(double* out, int iters, double av) {
    extern __shared__ double sm[];
    constexpr int KC = 16, LD = 128;
    for (int i = threadIdx.x; i < 2 * KC * LD; i += blockDim.x) sm[i] = av + (i % 97) * 1e-4;
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // rows tx*2 + 32*i (+0,1), i = 0..3 ; cols ty*2 + 32*j (+0,1)
    const double2* As = reinterpret_cast<const double2*>(sm) + tx;
    const double2* Bs = reinterpret_cast<const double2*>(sm + KC * LD) + ty;
    double c[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) c[i][j] = i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int k = 0; k < KC; ++k) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = As[k * (LD / 2) + 16 * i];
                a[2 * i] = v.x; a[2 * i + 1] = v.y;
                const double2 w = Bs[k * (LD / 2) + 16 * i];
                b[2 * i] = w.x; b[2 * i + 1] = w.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) c[i][j] = fma(a[i], b[j], c[i][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

// warps with (warp & 1) == 0 run DMMA, the others DFMA chains (16 independent accumulators per thread)
template <int DMMA_WARPS_OF_4>
__global__ void __launch_bounds__(512) mix_kernel(double* out, int iters_mma, int iters_fma, double av, double bv) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if ((warp & 3) < DMMA_WARPS_OF_4) {
        double c[8][2];
        double a = av + threadIdx.x * 1e-6, b = bv;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = 0; }
        for (int it = 0; it < iters_mma; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    } else {
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters_fma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], av, bv);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
    }
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
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    const double A = 1.0000001, B = 1e-9;

#define DMMA_SAME(NACC, threads, iters, tag) { \
        float ms = time_ms([&] { dmma_same<NACC, threads><<<sms, threads>>>(out, iters, A, B); }, 3); \
        double fl = 512.0 * NACC * (double)(iters) * (threads / 32) * sms; \
        printf("{\"test\":\"dmma_same\",\"run\":\"%s\",\"nacc\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", tag, NACC, threads / 32, ms, fl / ms * 1e-9); }
    // short (~1-2 ms) and long (~30-60 ms) runs: separates issue effects from clock/power effects
    DMMA_SAME(1, 256, 80000, "short"); DMMA_SAME(2, 256, 40000, "short"); DMMA_SAME(4, 256, 20000, "short");
    DMMA_SAME(8, 256, 10000, "short"); DMMA_SAME(16, 256, 5000, "short"); DMMA_SAME(32, 256, 2500, "short");
    DMMA_SAME(4, 256, 600000, "long"); DMMA_SAME(16, 256, 150000, "long"); DMMA_SAME(32, 256, 75000, "long");
    DMMA_SAME(1, 128, 80000, "short"); DMMA_SAME(2, 128, 40000, "short"); DMMA_SAME(4, 128, 20000, "short");
    DMMA_SAME(8, 128, 10000, "short"); DMMA_SAME(16, 128, 5000, "short"); DMMA_SAME(32, 128, 2500, "short");
    DMMA_SAME(4, 512, 10000, "short"); DMMA_SAME(8, 512, 5000, "short"); DMMA_SAME(16, 512, 2500, "short");
    DMMA_SAME(2, 1024, 10000, "short"); DMMA_SAME(4, 1024, 5000, "short"); DMMA_SAME(8, 1024, 2500, "short");

#define DMMA_TILE(MI, NI, threads, iters, tag) { \
        float ms = time_ms([&] { dmma_tile<MI, NI, threads><<<sms, threads>>>(out, iters, A, B); }, 3); \
        double fl = 512.0 * MI * NI * (double)(iters) * (threads / 32) * sms; \
        printf("{\"test\":\"dmma_tile\",\"run\":\"%s\",\"mi\":%d,\"ni\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", tag, MI, NI, threads / 32, ms, fl / ms * 1e-9); }
    DMMA_TILE(4, 8, 256, 2500, "short"); DMMA_TILE(4, 8, 256, 75000, "long");
    DMMA_TILE(2, 4, 256, 10000, "short"); DMMA_TILE(4, 4, 256, 5000, "short"); DMMA_TILE(2, 8, 256, 5000, "short");
    DMMA_TILE(4, 4, 512, 2500, "short"); DMMA_TILE(2, 4, 512, 5000, "short"); DMMA_TILE(2, 2, 1024, 5000, "short");
    DMMA_TILE(4, 8, 128, 2500, "short");

#define DMMA_SMEM(MI, NI, threads, iters, tag) { \
        CK(cudaFuncSetAttribute(dmma_smem<MI, NI, threads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 132 * 8)); \
        float ms = time_ms([&] { dmma_smem<MI, NI, threads><<<sms, threads, 2 * 16 * 132 * 8>>>(out, iters, A); }, 3); \
        double fl = 512.0 * MI * NI * 4.0 * (double)(iters) * (threads / 32) * sms; \
        printf("{\"test\":\"dmma_smem\",\"run\":\"%s\",\"mi\":%d,\"ni\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", tag, MI, NI, threads / 32, ms, fl / ms * 1e-9); }
    DMMA_SMEM(4, 8, 256, 800, "short"); DMMA_SMEM(4, 8, 256, 20000, "long");
    DMMA_SMEM(4, 4, 512, 800, "short"); DMMA_SMEM(2, 4, 512, 1600, "short");
    DMMA_SMEM(4, 4, 256, 1600, "short"); DMMA_SMEM(2, 8, 256, 1600, "short");

    {
        CK(cudaFuncSetAttribute(dfma_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 128 * 8));
        for (int bps : {1, 2}) {
            for (int iters : {400, 10000}) {
                float ms = time_ms([&] { dfma_tile<<<sms * bps, 256, 2 * 16 * 128 * 8>>>(out, iters, A); }, 3);
                double fl = 2.0 * 64 * 16 * (double)iters * 256 * sms * bps;
                printf("{\"test\":\"dfma_tile\",\"run\":\"%s\",\"blocks_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", iters > 1000 ? "long" : "short", bps, ms, fl / ms * 1e-9);
            }
        }
    }
    {
        // DMMA share of warps: 0/4 (pure DFMA), 1/4, 2/4, 3/4, 4/4 (pure DMMA). Iteration counts chosen so both kinds of
        // warp would take about equally long at the 37 TF/s peak if the pipes were shared fairly.
        const int threads = 512;  // 16 warps per SM, 4 per SMSP
        // per-warp work: DMMA iteration = 8 DMMA = 4096 flop ; DFMA iteration = 16 * 32 * 2 = 1024 flop
#define MIX(NM, im, ifm) { \
            float ms = time_ms([&] { mix_kernel<NM><<<sms, threads>>>(out, im, ifm, A, B); }, 3); \
            double wm = (threads / 32) * NM / 4.0, wf = (threads / 32) * (4 - NM) / 4.0; \
            double flm = 4096.0 * (im) * wm * sms, flf = 1024.0 * (ifm) * wf * sms; \
            printf("{\"test\":\"mix\",\"dmma_warps_of_4\":%d,\"iters_mma\":%d,\"iters_fma\":%d,\"ms\":%.3f,\"tflops_dmma\":%.3f,\"tflops_dfma\":%.3f,\"tflops_total\":%.3f}\n", \
                   NM, im, ifm, ms, flm / ms * 1e-9, flf / ms * 1e-9, (flm + flf) / ms * 1e-9); }
        MIX(0, 0, 40000); MIX(4, 10000, 0);
        MIX(2, 20000, 80000); MIX(2, 20000, 40000); MIX(2, 10000, 80000);
        MIX(1, 40000, 40000); MIX(3, 10000, 120000);
    }
    return 0;
}
