// covariance.cuh — host entry of the Gram / cross-covariance writer (pair-tile engine + CovWriteEpi, cov.cu).
#pragma once

#include "kernel_eval.cuh"
#include "model.cuh"
#include "pair_tiles.cuh"

namespace fgp {

// out[r + c*ld] = k(row point r, column point c) over the padded extents of `pa`; symmetric => lower triangle only,
// noise2 added on the diagonal, identity on the padding (make_cholesky_cov_matrix algebra/mod.rs:67-79 /
// make_covariance_matrix algebra/mod.rs:41-54).
void write_covariance(fgp_model* m, const KernelTraits& kt, const fgp_kernel_desc* kd, const PairArgs& pa, double* out,
                      int64_t ld, int64_t valid_rows, int64_t valid_cols, double noise2);

}  // namespace fgp
