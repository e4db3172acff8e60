// cov.cu — instantiations of the pair-tile engine that write covariance matrices (contract in covariance.cuh).
#include "covariance.cuh"

#include <cmath>

namespace fgp {

// ---- Gram / cross-covariance launch -----------------------------------------------------------------------------
template <int KIND, int MODE>
void launch_cov(const PairArgs& pa, const CovWriteEpi<KIND>& epi, cudaStream_t st) {
    pair_tile_kernel<MODE, CovWriteEpi<KIND>><<<pair_grid(pa), 256, 0, st>>>(pa, epi);
}

void write_covariance(fgp_model* m, const KernelTraits& kt, const fgp_kernel_desc* kd, const PairArgs& pa, double* out,
                      int64_t ld, int64_t valid_rows, int64_t valid_cols, double noise2) {
    const DevKernel dk = to_dev(kd);
    const int mode = (kt.need_d2 ? PAIR_D2 : 0) | (kt.need_dot ? PAIR_DOT : 0);
    const LaunchCtx lc = m->ctx();
    ProfScope ps(lc, PROF_PAIR, (double)pa.rows * pa.cols * (pa.symmetric ? 0.5 : 1.0) * 2.0 * pa.dp);
    if (kt.kind == KIND_SQEXP) {
        const double ls = dk.param[0];
        CovWriteEpi<KIND_SQEXP> e{dk, out, ld, valid_rows, valid_cols, pa.symmetric, noise2,
                                  -1.0 / (2.0 * ls * ls), std::fabs(dk.param[1]), 0.0};
        launch_cov<KIND_SQEXP, PAIR_D2>(pa, e, m->st);
    } else if (kt.kind == KIND_MATERN2) {
        const double l = std::fabs(dk.param[0]);
        CovWriteEpi<KIND_MATERN2> e{dk, out, ld, valid_rows, valid_cols, pa.symmetric, noise2,
                                    std::sqrt(5.0) / l, std::fabs(dk.param[1]), 5.0 / (3.0 * l * l)};
        launch_cov<KIND_MATERN2, PAIR_D2>(pa, e, m->st);
    } else {
        CovWriteEpi<KIND_GENERIC> e{dk, out, ld, valid_rows, valid_cols, pa.symmetric, noise2, 0.0, 0.0, 0.0};
        if (mode == PAIR_D2) launch_cov<KIND_GENERIC, PAIR_D2>(pa, e, m->st);
        else if (mode == PAIR_DOT) launch_cov<KIND_GENERIC, PAIR_DOT>(pa, e, m->st);
        else launch_cov<KIND_GENERIC, PAIR_BOTH>(pa, e, m->st);
    }
    m->launches += 1;
}

}  // namespace fgp
