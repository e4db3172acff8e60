// ozaki.cuh — the trailing update of the blocked Cholesky on the 5th-generation tensor cores (tcgen05 / TMEM).
//
//        C(lower) -= P P^T        P = the solved panel (rows x K f64, K <= 512), C = the trailing matrix behind it
//
// replaces, like gemm_nt.cuh, the axpy chain of nalgebra's Cholesky::new_internal (called at src/algebra/mod.rs:83,90).
// `tcgen05.mma` has no f64 kind, so the product is formed EXACTLY out of integer products (the "Ozaki
// scheme"): every row of P is scaled by a power of two (its own exponent e_r: max |p| < 2^e_r) and cut into
// OZ_SLICES = 8 balanced base-128 digits,
//        p[r][k] = 2^(e_r - 55) * sum_i d_i[r][k] * 128^(7-i),        d_i in [-65, 65]  (int8),
// which keeps 55 bits below the row's leading bit (fp64 itself keeps 53 below each ELEMENT's leading bit).  Then
//        (P P^T)[r][c] = 2^(e_r + e_c - 110) * sum_{i,j} 128^(14-i-j) * (D_i D_j^T)[r][c],      D_i D_j^T : int8 x int8 -> int32,
// and `tcgen05.mma.kind::i8` accumulates each D_i D_j^T over K <= 512 in int32 TMEM accumulators without any rounding
// (|sum| <= 8 * 512 * 65 * 64 < 2^31).  Slice pairs with i + j > 7 are dropped: they carry less than 2^-53 of
// max_row * max_col per term — the size of the rounding of one f64 multiply-add.  The 36 kept products are grouped by
// g = i + j (equal weight 128^(14-g)): one accumulator per group, four groups per pass through K (TMEM holds four
// 128x128 int32 accumulators), two passes per tile: g = 0..3 (10 products) and g = 4..7 (26 products).  After each pass
// the four accumulators are combined exactly (64-bit integers, then f64: every partial sum < 2^47), scaled by the row /
// column powers of two (exact) and added to C by TMA reduce-adds (f64 add at the L2) — two roundings per element and
// launch, against 512 in the f64 DMMA kernel.
//
// Data layout.  `ozaki_slice_launch` writes the digits in the order the tensor core reads them, so the GEMM kernel moves
// them with 1-D bulk copies and no swizzle: for row tile T (128 rows), k-step s (32 contraction bytes = one tcgen05.mma) and
// slice i the 4 KB block at ((T * KS + s) * 8 + i) * 4096 holds the 128 x 32 digits as UMMA K-major "core matrices"
// (8 rows x 16 bytes, contiguous 128 B): byte offset = (k / 16) * 2048 + (r / 8) * 128 + (r % 8) * 16 + k % 16, i.e. the
// shared-memory descriptor's leading-dimension byte offset (between the two k halves) is 2048 and its stride byte offset
// (between 8-row groups) is 128.  One stage of the operand ring is the 8 slices of the tile's row block (32 KB) and of its
// column block (32 KB) for one k-step; pass 0 only fetches slices 0..3 of each.
//
// Kernel roles (320 threads, one CTA per SM): warp 0 = TMEM allocation + the bulk copies, warp 1 = the MMAs (36 x K/32 per
// tile) and the commits that release stages / publish accumulators — both warps walk their loops as a whole (uniform
// control flow and descriptor arithmetic) and one elected lane issues; warps 2..9 = epilogue: tcgen05.ld of the warp's 32
// TMEM lanes (= 32 tile rows) x 64 columns into registers (one exact f64 per element), release of the accumulators, then
// scaling and TMA reduce-adds into C from the warp's own 32 x 16 staging image while the next pass's MMAs already run.
#pragma once

#include "gemm_nt.cuh"

namespace fgp {

constexpr int OZ_SLICES = 8;
constexpr int OZ_KSTEP = 32;                          // contraction length of one tcgen05.mma kind::i8
constexpr int OZ_BLOCK_BYTES = 128 * OZ_KSTEP;        // one slice of one row tile for one k-step
constexpr int OZ_STAGES = 3;
constexpr int OZ_PART_BYTES = OZ_SLICES * OZ_BLOCK_BYTES;                 // 32 KB: row (or column) block of one stage
constexpr int OZ_STAGE_BYTES = 2 * OZ_PART_BYTES;                          // 64 KB
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_STAGING_BYTES = OZ_EPI_WARPS * 32 * 16 * 8;               // one 32 x 16 f64 image per epilogue warp
constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + OZ_STAGING_BYTES + 1024;
constexpr int OZ_THREADS = 32 * (2 + OZ_EPI_WARPS);
constexpr uint32_t OZ_LBO = 16, OZ_SBO = 256;   // SWIZZLE_32B, K-major: 8-row groups are 256 bytes apart; LBO unused

// bytes of the digit blob / doubles of the scale vector of a rows x K panel (rows % 128 == 0, K % 32 == 0)
inline size_t ozaki_slice_bytes(int64_t rows, int K) { return (size_t)rows * (size_t)K * OZ_SLICES; }

cudaError_t ozaki_prepare();
void ozaki_set_experiment(int flags);  // measurement switches of the update kernel (0 = production)
// digits + row scales (2^(e_r - 30), 0 for an all-zero row, NaN when the row holds a non-finite value) of P (column-major, ld)
void ozaki_slice_launch(const double* P, int64_t ld, int64_t rows, int K, int8_t* digits, double* scale, const LaunchCtx& ctx);
// C += alpha * P Q^T on the tiles GemmArgs describes (C, ldc, M, N, K, alpha = +-1, lower, row_skip, grp, stride, k_from_tile,
// k_tile0; beta = 1 implied: a plain product needs C zeroed first); tile row ti reads row tile ti of (digitsA, scaleA), tile
// column tj row tile tj of (digitsB, scaleB).  K <= 32768 keeps the int32 accumulators exact (8 K 65 64 < 2^31).
// tiles_per_cta <= 0: default. Returns the number of tiles launched.
int64_t ozaki_update_launch(const GemmArgs& g, const int8_t* digitsA, const double* scaleA, const int8_t* digitsB,
                            const double* scaleB, int tiles_per_cta, const LaunchCtx& ctx, uint32_t lbo = OZ_LBO,
                            uint32_t sbo = OZ_SBO);

}  // namespace fgp
