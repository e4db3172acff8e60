// sharded.cuh — multi-GPU (one process per GPU) block-column Cholesky: 1-D block-cyclic ownership of the 512-wide
// panels, NCCL broadcast of each factored panel over NVLink, one-panel look-ahead.  The reference has
// no counterpart (single-threaded crate); the entry points are declared in include/fgp.h (fgp_comm_*, fgp_fit_sharded).
#pragma once

#include <algorithm>
#include <utility>
#include <vector>

#include "nccl_dyn.cuh"

#include "kernel_eval.cuh"
#include "model.cuh"
#include "potrf.cuh"

struct fgp_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    fgp::DevBuf pbuf[2];                     // contiguous panel buffers: rows [J*128, np) x panel width, ld = rows
    cudaStream_t st_comm = nullptr;          // packs and broadcasts the panel slab by slab while the panel stream factors on
    cudaStream_t st_copy = nullptr;          // factor_sharded_pipe: copies of the received panel pieces into L (high priority: a buffer is free again only after its copy, and a low-priority copy would queue behind every pending CTA of the trailing update)
    cudaEvent_t ev_col = nullptr;            // a block column of the panel is final (recorded on the panel stream)
    cudaEvent_t ev_bcast = nullptr;          // panel J has arrived (recorded on the comm stream)
    cudaEvent_t ev_trail[2] = {nullptr, nullptr};  // trailing update with panel J done (main stream), index J & 1
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};   // this rank's copy of panel buffer J & 1 back into L done (side stream)
    std::vector<cudaEvent_t> ev_pipe;        // factor_sharded_pipe: [panel parity][solve | bcast | sliced | look-ahead][piece]
    double bcast_bytes = 0.0;                // bytes this rank sent or received in the last sharded factorisation
};

namespace fgp {

// panel p (PANEL_TILES block columns) is owned by rank p % nranks
inline int shard_owner(int64_t panel, int nranks) { return (int)(panel % nranks); }

// row pieces of the `below` rows under a panel's diagonal block in factor_sharded_pipe (multiples of 128 rows): the next panel's
// diagonal-block rows (512), the rows of the panel after it (<= 512), then the rest in equal pieces of at most pipe_rows rows.
// Appends (first row, height) pairs; below >= 1024.
inline void shard_pieces(int64_t below, int64_t pipe_rows, std::vector<std::pair<int64_t, int64_t>>& out) {
    int64_t r0 = 0;
    auto push = [&](int64_t h) {
        out.push_back({r0, h});
        r0 += h;
    };
    push(512);
    if (below - r0 > 0) push(std::min<int64_t>(512, below - r0));
    const int64_t rest_tiles = (below - r0) / 128;
    if (rest_tiles > 0) {
        const int64_t k = (rest_tiles * 128 + pipe_rows - 1) / pipe_rows;
        for (int64_t i = 0; i < k; ++i) push((rest_tiles / k + (i < rest_tiles % k ? 1 : 0)) * 128);
    }
}

// panel buffers for the current problem size (rank-local allocation; entry points call it BEFORE the cross-rank status exchange)
int reserve_sharded(fgp_model* m);

// Gram assembly of the owned panels + the sharded factorisation; on return every rank holds the complete factor in m->L.
int factor_sharded(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps);
// the same on the head schedule (one potrf_head_kernel per panel; bit-identical to the single-GPU head schedule)
int factor_sharded_head(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps,
                        const PotrfWork& w);

// the head schedule with the panel travelling in row pieces: solve, broadcast, digit slicing and the next owner's look-ahead
// overlap piece by piece, the next head starts after the first 4 MB (sharded.cu); same arithmetic, bit-identical factor
int factor_sharded_pipe(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps,
                        const PotrfWork& w);

}  // namespace fgp
