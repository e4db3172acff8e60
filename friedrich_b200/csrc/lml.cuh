// lml.cuh — gradient of the log marginal likelihood and the bandwidth heuristic, on the device.
//
// Reference (src/gaussian_process/optimizer.rs:24-60 unscaled, :159-203 scaled):
//     cov_inv = covmat_cholesky.inverse()                       (2 n^3 flops of scalar triangular solves as executed)
//     alpha   = cov_inv * y ;  scale = y.alpha / n              (scaled variant)
//     G_p     = make_gradient_covariance_matrices               (algebra/mod.rs:129-155: P dense n x n matrices)
//     grad_p  = ( alpha^T G_p alpha [/ scale]  -  tr(cov_inv G_p) ) / 2
//     grad_noise = noise * (alpha.alpha - tr cov_inv)           (unscaled variant only, optimizer.rs:54-57)
//
// Device path:  U = L^-T by the tensor-pipe forward solve on the identity (trsm.cuh, n^3/3), K^-1 = U U^T on the lower
// triangle (gemm_nt with k_from_tile, n^3/3), then ONE pass of the pair-tile engine over the lower triangle that
// re-evaluates the kernel gradients from X on the fly and reduces both  sum K^-1[r,c] G_p[r,c]  and
// sum alpha_r alpha_c G_p[r,c]  — the P gradient matrices are never materialised (the reference keeps all of them).
// Reductions are deterministic: per-CTA partials in a fixed tree, then one block sums the partials in order.
#pragma once

#include "gemm_nt.cuh"
#include "kernel_eval.cuh"
#include "model.cuh"
#include "pair_tiles.cuh"
#include "trsm.cuh"
#include "vector_kernels.cuh"

namespace fgp {

// defined in lml.cu
int lml_gradient_device(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int scaled,
                        double* scale_out, double* grads);
// collective: every rank of the model's communicator (all hold the full factor); same result on every rank
int lml_gradient_sharded_device(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int scaled,
                                double* scale_out, double* grads);
int mean_pair_distance_device(fgp_model* m, double* out);
// A[i + i*ld] = 1 for i < n
void launch_set_identity(double* A, int64_t ld, int64_t n, cudaStream_t st);

// LinearPrior::fit on the resident (centred) training inputs; y_host = the ORIGINAL outputs (n values)
int linear_prior_fit_device(fgp_model* m, const double* y_host, double* weights, double* intercept);

}  // namespace fgp
