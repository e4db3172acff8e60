// common.cuh — shared device helpers for libfgp_sm100 (sm_100a only).
//
// fp64 tensor path: there is no tcgen05 `.kind::f64` (ptxas rejects it); the fp64 tensor pipe of
// B200 is reached with `mma.sync.aligned.m8n8k4.f64` -> SASS DMMA.8x8x4.  Tiles move through the TMA engine: tensor copies
// (`cp.async.bulk.tensor.2d` -> UTMALDG / UTMASTG, and `cp.reduce.async.bulk.tensor` -> UTMAREDG for C += tile at the L2),
// 1-D bulk copies (`cp.async.bulk` -> UBLKCP) and L2 prefetches, completing on mbarriers / bulk groups.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace fgp {

constexpr int TILE = 128;  // every matrix on the device is padded to a multiple of TILE ("layout")

__host__ __device__ inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// Function attributes (opt-in shared memory sizes) are per device: one-time setup is tracked per CUDA device, so that a
// process holding models on several GPUs configures each of them. Returns the slot of the current device.
inline bool* per_device_flag(bool (&flags)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    return &flags[(dev >= 0 && dev < 64) ? dev : 0];
}

// ---------------------------------------------------------------------------------------------------------------
// Launch context: the stream plus an optional per-kernel-class profiler (CUDA events around each launch on the
// launching stream; bench.py's roofline figures come from it, include/fgp.h fgp_set_profiling).
enum ProfClass { PROF_GEMM = 0, PROF_POTRF_DIAG = 1, PROF_PAIR = 2, PROF_OTHER = 3, PROF_TCGEN05 = 4, PROF_SLICE = 5, PROF_NCLASS = 6 };

struct Profiler {
    struct Rec { int cls; double flops; cudaEvent_t e0, e1; };
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<Rec> recs;
    double ms[PROF_NCLASS] = {};
    double flops[PROF_NCLASS] = {};
    int64_t count[PROF_NCLASS] = {};
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void reset() {
        used = 0;
        recs.clear();
        for (int i = 0; i < PROF_NCLASS; ++i) { ms[i] = 0; flops[i] = 0; count[i] = 0; }
    }
    // call after the stream has been synchronised
    void collect() {
        for (const Rec& r : recs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) ms[r.cls] += t;
            flops[r.cls] += r.flops;
            count[r.cls] += 1;
        }
        recs.clear();
        used = 0;
    }
    void destroy() {
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
        pool.clear();
    }
};

struct LaunchCtx {
    cudaStream_t st = nullptr;
    Profiler* prof = nullptr;  // null: no per-launch events
};

struct ProfScope {
    const LaunchCtx& ctx;
    Profiler::Rec rec;
    ProfScope(const LaunchCtx& c, int cls, double flops) : ctx(c) {
        if (ctx.prof) {
            rec = {cls, flops, ctx.prof->get(), ctx.prof->get()};
            cudaEventRecord(rec.e0, ctx.st);
        }
    }
    ~ProfScope() {
        if (ctx.prof) {
            cudaEventRecord(rec.e1, ctx.st);
            ctx.prof->recs.push_back(rec);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// DMMA: D(8x8) += A(8x4, row) * B(4x8, col).  Lane l: g = l >> 2, t = l & 3.
//   a  = A[g][t]      b = B[t][g]      c0 = C[g][2t]   c1 = C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (1-D) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy through the TMA engine; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// global -> shared 2-D tile through the TMA engine (cp.async.bulk.tensor -> SASS UTMALDG): the box described by the tensor
// map at coordinates (c0 = index along the contiguous dimension, c1 = column); out-of-range elements arrive as zeros and
// the mbarrier is always credited with the full box size
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// shared -> global 2-D tile through the TMA engine (SASS UTMASTG / UTMAREDG), bulk-group completion:
//   tma_store_2d   C[box] = smem            tma_reduce_add_2d   C[box] += smem  (element-wise f64 add performed at the L2)
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tmap), "r"(c0),
                 "r"(c1), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, int c0, int c1, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(tmap),
                 "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the shared-memory source of every committed bulk group has been read (the CTA may exit / reuse the buffer)
__device__ __forceinline__ void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// asynchronous L2 prefetch of `bytes` (multiple of 16) starting at a 16-byte aligned global address
__device__ __forceinline__ void l2_prefetch(const void* gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace fgp
