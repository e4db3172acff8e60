// sharded.cu — see sharded.cuh.  Schedule per rank r of P, for panel p = 0 .. NP-1 (owner = p % P):
//
//   panel stream (high priority)                                   main stream
//   ---------------------------------------------------------      ------------------------------------------------
//   wait: trailing update p-2 done (buffer p&1 free, my panel
//         p has every update up to p-2)
//   owner: look-ahead update of panel p by panel p-1 (buffer)
//          factor panel p (update / diagonal tile / panel solve)
//          pack rows >= p*512 of the panel into buffer p&1
//   ncclBroadcast(buffer p&1, root = owner)            ---->      wait: panel p arrived
//                                                                  update every owned panel c > p (except c = p+1:
//                                                                  that is the owner's look-ahead above) from the buffer
//                                                                  non-owners: unpack the buffer into their copy of L
//
// so the factorisation + broadcast of panel p+1 overlaps the trailing updates by panel p on all ranks, and when the loop
// ends every rank holds the whole factor (replicated L: alpha solve and predict run locally, queries shard trivially).
//
// Three schedules share this distribution: factor_sharded (per-block-column chain of round 1, FGP_OPT_HEAD = 0),
// factor_sharded_head (one head launch per panel, the panel broadcast in one piece, FGP_OPT_SHARD_PIPE = 0) and the default
// factor_sharded_pipe at the end of this file (the panel travels in row pieces; solve / broadcast / slicing / look-ahead
// overlap piece by piece).  All three produce the factor of the single-GPU schedule they mirror bit for bit.
#include "ozaki.cuh"
#include "sharded.cuh"

#include <cstdlib>
#include <vector>

#include <algorithm>
#include <functional>
#include <string>

#include "covariance.cuh"
#include "gemm_nt.cuh"

namespace fgp {

#define NC(m, call)                                                                                                \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess)                                                                                    \
            return fgp::fail(m, FGP_ERR_COMM, std::string(#call) + ": " + nccl->GetErrorString(r__));              \
    } while (0)

// info holds 0 or the 1-based failing column of THIS rank; the first failure anywhere is the minimum over the non-zero ones
static __global__ void info_encode_kernel(int* info, int decode) {
    if (!decode) *info = (*info == 0) ? 0x7fffffff : *info;
    else *info = (*info == 0x7fffffff) ? 0 : *info;
}

// owned block columns [c0, c1) (rows >= c0) -= B B^T restricted to them, B = packed panel buffer holding rows >= J*128
static void update_from_buffer(fgp_model* m, const double* buf, int64_t J, int64_t Jend, int64_t c0, int64_t c1,
                               const LaunchCtx& c, PotrfCounters* cnt) {
    const int64_t rows = m->np - J * TILE;
    GemmArgs g{};
    g.C = m->L.p + c0 * TILE + c0 * TILE * m->cap; g.ldc = m->cap;
    g.A = buf + (c0 - J) * TILE; g.lda = rows;
    g.B = g.A; g.ldb = rows;
    g.M = (int)(m->np - c0 * TILE); g.N = (int)((c1 - c0) * TILE); g.K = (int)((Jend - J) * TILE);
    g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.k_from_tile = 0;
    cnt->launches += gemm_nt_launch(g, c) > 0;
}

int reserve_sharded(fgp_model* m) {
    fgp_comm* cm = m->comm;
    const int64_t W = (int64_t)panel_tiles(m->np, cm->nranks) * TILE;
    CU(m, cm->pbuf[0].reserve((size_t)m->np * W));
    CU(m, cm->pbuf[1].reserve((size_t)m->np * W));
    return FGP_OK;
}

int factor_sharded(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps) {
    fgp_comm* cm = m->comm;
    const int P = cm->nranks, r = cm->rank;
    const NcclApi* nccl = nccl_api();
    if (P > 1 && !nccl) return fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
    const int64_t np = m->np, nb = np / TILE, PANEL_TILES = panel_tiles(np, P), NP = (nb + PANEL_TILES - 1) / PANEL_TILES;
    FGP_TRY(reserve_sharded(m));  // no-op when the entry point has already done it (before the status exchange)
    CU(m, cudaMemsetAsync(m->info_d, 0, sizeof(int), m->st));
    cm->bcast_bytes = 0.0;

    // ---- Gram: only the block columns this rank owns (algebra/mod.rs:67-79 restricted to them) ---------------------
    for (int64_t p = r; p < NP; p += P) {
        const int64_t J = p * PANEL_TILES, Jend = std::min<int64_t>(J + PANEL_TILES, nb);
        PairArgs pa{};
        pa.xa_c = pa.xb_c = m->xc.p;
        pa.xa_r = pa.xb_r = m->xr.p;
        pa.na = pa.nb = m->nc.p;
        pa.dp = (int)m->dp;
        pa.rows = pa.cols = np;
        pa.row_tile0 = (int)J;
        pa.col_tile0 = (int)(J * TILE / PAIR_TN);
        pa.col_tiles = (int)((Jend - J) * TILE / PAIR_TN);
        pa.symmetric = 1;
        write_covariance(m, kt, kd, pa, m->L.p, m->cap, m->n, m->n, noise * noise);
    }

    const LaunchCtx mc = m->ctx();
    LaunchCtx pc = mc;
    pc.st = m->st2;
    PotrfCounters cnt;
    CU(m, cudaEventRecord(m->evA, m->st));
    CU(m, cudaStreamWaitEvent(m->st2, m->evA, 0));
    for (int64_t p = 0; p < NP; ++p) {
        const int64_t J = p * PANEL_TILES, Jend = std::min<int64_t>(J + PANEL_TILES, nb);
        const int64_t rows = np - J * TILE, w = (Jend - J) * TILE;
        const int owner = shard_owner(p, P);
        double* buf = cm->pbuf[p & 1].p;
        if (p >= 2) {  // the trailing update with panel p-2 has left this buffer and reached this panel's columns
            CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[p & 1], 0));
            CU(m, cudaStreamWaitEvent(cm->st_comm, cm->ev_trail[p & 1], 0));
        }
        // The panel travels slab by slab (one 128-column block each): as soon as the owner has finalised a block column it
        // is packed into the contiguous panel buffer and broadcast on the comm stream, while the panel stream factors the
        // next block column — only the last slab's transfer is on the critical path.
        int rc_ship = FGP_OK;
        auto ship = [&](int64_t j, bool is_owner) {
            double* slab = buf + (j - J) * TILE * rows;
            if (is_owner) {
                cudaEventRecord(cm->ev_col, m->st2);
                cudaStreamWaitEvent(cm->st_comm, cm->ev_col, 0);
                cudaMemcpy2DAsync(slab, rows * sizeof(double), m->L.p + J * TILE + j * TILE * m->cap, m->cap * sizeof(double),
                                  rows * sizeof(double), TILE, cudaMemcpyDeviceToDevice, cm->st_comm);
            }
            if (P > 1) {
                // profiled (class "other") only on the owner: there the duration is the send itself, elsewhere it includes
                // the wait for the owner's factorisation; the "flops" slot carries the bytes
                LaunchCtx bc = mc;
                bc.st = cm->st_comm;
                if (!is_owner) bc.prof = nullptr;
                ProfScope ps(bc, PROF_OTHER, (double)rows * TILE * sizeof(double));
                if (nccl->Broadcast(slab, slab, (size_t)rows * TILE, ncclDouble, owner, cm->comm, cm->st_comm) != ncclSuccess)
                    rc_ship = FGP_ERR_COMM;
                cm->bcast_bytes += (double)rows * TILE * sizeof(double);
            }
        };
        if (owner == r) {
            const std::function<void(int64_t)> on_col = [&](int64_t j) { ship(j, true); };
            if (p >= 1 && Jend - J > 1) {
                // look-ahead update by the previous panel: block column 0 here (its factorisation follows at once), block
                // columns 1.. on the side stream, joined before the panel stream touches block column 1
                const int64_t Jp = (p - 1) * PANEL_TILES;
                LaunchCtx sc = mc;
                sc.st = m->st3;
                CU(m, cudaEventRecord(m->evC, m->st2));
                CU(m, cudaStreamWaitEvent(m->st3, m->evC, 0));
                update_from_buffer(m, cm->pbuf[(p - 1) & 1].p, Jp, J, J, J + 1, pc, &cnt);
                update_from_buffer(m, cm->pbuf[(p - 1) & 1].p, Jp, J, J + 1, Jend, sc, &cnt);
                CU(m, cudaEventRecord(m->evC, m->st3));
                const PanelSide ps{m->st3, m->evD, m->evC};
                factor_panel(m->L.p, m->cap, np, J, Jend, m->inv.p, m->invT.p, has_eps, eps, m->info_d, pc, &cnt, &ps, &on_col);
            } else {
                if (p >= 1) {
                    const int64_t Jp = (p - 1) * PANEL_TILES;
                    update_from_buffer(m, cm->pbuf[(p - 1) & 1].p, Jp, J, J, Jend, pc, &cnt);
                }
                const PanelSide ps{m->st3, m->evD, m->evC};  // evC: at most an old, completed record (nothing ran on st3 for it)
                factor_panel(m->L.p, m->cap, np, J, Jend, m->inv.p, m->invT.p, has_eps, eps, m->info_d, pc, &cnt, &ps, &on_col);
            }
        } else {
            for (int64_t j = J; j < Jend; ++j) ship(j, false);
        }
        if (rc_ship != FGP_OK) return fail(m, FGP_ERR_COMM, "ncclBroadcast of a panel slab failed");
        CU(m, cudaEventRecord(cm->ev_bcast, cm->st_comm));
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_bcast, 0));  // the next owner's look-ahead update reads the buffer on st2
        CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
        {
            // every owned panel c > p, c != p+1 (that one is the owner's look-ahead on the panel stream next iteration),
            // in ONE launch: the owned panels are groups of PANEL_TILES tile columns, P*PANEL_TILES tile columns apart
            int64_t c_first = p + 1 + ((r - (p + 1)) % P + P) % P;
            if (c_first == p + 1) c_first += P;
            if (c_first < NP) {
                const int64_t c0 = c_first * PANEL_TILES;
                int64_t ncols = 0;
                for (int64_t c = c_first; c < NP; c += P) ncols += std::min<int64_t>(PANEL_TILES, nb - c * PANEL_TILES);
                GemmArgs g{};
                g.C = m->L.p + c0 * TILE + c0 * TILE * m->cap; g.ldc = m->cap;
                g.A = buf + (c0 - J) * TILE; g.lda = rows;
                g.B = g.A; g.ldb = rows;
                g.M = (int)(np - c0 * TILE); g.N = (int)(ncols * TILE); g.K = (int)w;
                g.alpha = -1.0; g.beta_one = 1; g.lower = 1; g.k_from_tile = 0;
                g.grp = (int)PANEL_TILES; g.stride = (int)(P * PANEL_TILES);
                cnt.launches += gemm_nt_launch(g, mc) > 0;
            }
        }
        if (owner != r)
            CU(m, cudaMemcpy2DAsync(m->L.p + J * TILE + J * TILE * m->cap, m->cap * sizeof(double), buf,
                                    rows * sizeof(double), rows * sizeof(double), w, cudaMemcpyDeviceToDevice, m->st));
        CU(m, cudaEventRecord(cm->ev_trail[p & 1], m->st));
    }
    // the inverted diagonal tiles live with their owners; every rank needs them for the triangular solves that follow
    if (P > 1) {
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[(NP - 1) & 1], 0));
        NC(m, nccl->GroupStart());
        for (int64_t p = 0; p < NP; ++p) {
            const int64_t J = p * PANEL_TILES, Jend = std::min<int64_t>(J + PANEL_TILES, nb);
            const size_t cnt_d = (size_t)(Jend - J) * TILE * TILE;
            NC(m, nccl->Broadcast(m->inv.p + J * TILE * TILE, m->inv.p + J * TILE * TILE, cnt_d, ncclDouble, shard_owner(p, P),
                                cm->comm, m->st2));
            NC(m, nccl->Broadcast(m->invT.p + J * TILE * TILE, m->invT.p + J * TILE * TILE, cnt_d, ncclDouble,
                                shard_owner(p, P), cm->comm, m->st2));
        }
        NC(m, nccl->GroupEnd());
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 0);
        NC(m, nccl->AllReduce(m->info_d, m->info_d, 1, ncclInt, ncclMin, cm->comm, m->st2));  // first failing column anywhere
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 1);
        CU(m, cudaEventRecord(cm->ev_bcast, m->st2));
        CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
    }
    m->launches += cnt.launches;
    return FGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same distribution on the HEAD schedule (potrf.cuh): per panel p the owner runs, on its panel stream,
//     diagonal block of p  -= (rows of p in panel p-1) (..)^T          gemm_nt lower, from the previous panel buffer
//     L11, W = L11^-1                                                   potrf_head_kernel (one launch)
//     rows below in p's columns -= (panel p-1 rows below)(rows of p)^T  gemm_nt, on the side stream beside the head
//     buffer rows below = A21 W^T                                       gemm_nt, K <= 512, out of place
//     buffer top = L11                                                  2-D copy
// then the whole panel buffer is broadcast; every rank applies it to the panels it owns behind p+1 in one launch on the main
// stream (panel p+1 is its owner's look-ahead above) and keeps a copy of the panel in its L.  The arithmetic per tile is
// that of the single-GPU head schedule, so the factors are bit-identical.
int factor_sharded_head(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps,
                        const PotrfWork& w) {
    fgp_comm* cm = m->comm;
    const int P = cm->nranks, r = cm->rank;
    const NcclApi* nccl = nccl_api();
    if (P > 1 && !nccl) return fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
    const int64_t np = m->np, nb = np / TILE, PT = HEAD_PANEL / TILE, NP = (nb + PT - 1) / PT;
    FGP_TRY(reserve_sharded(m));
    CU(m, cudaMemsetAsync(m->info_d, 0, sizeof(int), m->st));
    CU(m, cudaMemsetAsync(w.sync, 0, (size_t)NP * HEAD_SYNC_INTS * sizeof(int), m->st));
    cm->bcast_bytes = 0.0;
    for (int64_t p = r; p < NP; p += P) {  // Gram: only the block columns this rank owns (algebra/mod.rs:67-79 restricted to them)
        const int64_t J = p * PT, Jend = std::min<int64_t>(J + PT, nb);
        PairArgs pa{};
        pa.xa_c = pa.xb_c = m->xc.p;
        pa.xa_r = pa.xb_r = m->xr.p;
        pa.na = pa.nb = m->nc.p;
        pa.dp = (int)m->dp;
        pa.rows = pa.cols = np;
        pa.row_tile0 = (int)J;
        pa.col_tile0 = (int)(J * TILE / PAIR_TN);
        pa.col_tiles = (int)((Jend - J) * TILE / PAIR_TN);
        pa.symmetric = 1;
        write_covariance(m, kt, kd, pa, m->L.p, m->cap, m->n, m->n, noise * noise);
    }
    const LaunchCtx mc = m->ctx();
    LaunchCtx pc = mc, sc = mc;
    pc.st = m->st2;
    sc.st = m->st3;
    PotrfCounters cnt;
    CU(m, cudaEventRecord(m->evA, m->st));
    CU(m, cudaStreamWaitEvent(m->st2, m->evA, 0));
    auto gemm = [&](double* Cp, int64_t ldc, const double* Ap, int64_t lda, const double* Bp, int64_t ldb, int64_t M, int64_t N,
                    int64_t K, double alpha, int beta_one, int lower, int k_upto, const LaunchCtx& c) {
        GemmArgs g{};
        g.C = Cp; g.ldc = ldc;
        g.A = Ap; g.lda = lda;
        g.B = Bp; g.ldb = ldb;
        g.M = (int)M; g.N = (int)N; g.K = (int)K;
        g.alpha = alpha; g.beta_one = beta_one; g.lower = lower; g.k_upto_col = k_upto;
        cnt.launches += gemm_nt_launch(g, c) > 0;
    };
    bool copy_pending[2] = {false, false};
    // tiles per CTA of the main-stream updates when several GPUs take turns as panel owners: short-lived CTAs (one tile, 34 us) hand
    // SMs to the owner's high-priority chain kernels (head, panel solve) sooner than the wave-filling default (up to 4 tiles)
    static const int OZ_SHARDED_TPC = getenv("FGP_SHARD_TPC") ? atoi(getenv("FGP_SHARD_TPC")) : 1;
    // FGP_SHARD_TRACE=1: per-panel timeline of this rank (ms since the first record), printed by every rank at the end of the fit
    static const bool trace = getenv("FGP_SHARD_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](int64_t pnl, int what, cudaStream_t stm) {
        if (!trace) return;
        if (tev.empty()) {
            tev.resize((size_t)NP * 6 + 1);
            for (auto& e : tev) cudaEventCreate(&e);
            cudaEventRecord(tev.back(), m->st);
        }
        cudaEventRecord(tev[(size_t)pnl * 6 + what], stm);
    };
    // digit slices of panel q: kept per panel when the model holds a store for them (prepare_head_work), else one scratch image
    auto oz_dig = [&](int64_t q) { return w.oz_digits + ((w.oz_off_bytes && w.oz_off_bytes[q] >= 0) ? w.oz_off_bytes[q] : 0); };
    auto oz_sc = [&](int64_t q) { return w.oz_scale + ((w.oz_off_rows && w.oz_off_rows[q] >= 0) ? w.oz_off_rows[q] : 0); };
    for (int64_t p = 0; p < NP; ++p) {
        const int64_t J = p * PT, Jend = std::min<int64_t>(J + PT, nb);
        const int64_t rows = np - J * TILE, wc = (Jend - J) * TILE, below = rows - wc;
        const int owner = shard_owner(p, P);
        double* buf = cm->pbuf[p & 1].p;
        if (p >= 2) {  // the trailing update with panel p-2 has left this buffer and reached this panel's columns
            CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[p & 1], 0));
            CU(m, cudaStreamWaitEvent(cm->st_comm, cm->ev_trail[p & 1], 0));
            if (copy_pending[p & 1]) {  // ... and so has this rank's copy of it into L
                CU(m, cudaStreamWaitEvent(m->st2, cm->ev_copy[p & 1], 0));
                CU(m, cudaStreamWaitEvent(cm->st_comm, cm->ev_copy[p & 1], 0));
            }
        }
        if (owner == r) {
            double* Ajj = m->L.p + J * TILE * (m->cap + 1);
            double* A21 = m->L.p + Jend * TILE + J * TILE * m->cap;
            if (p >= 1) {  // look-ahead: this panel's columns get the previous panel's update here, not in the main-stream launch
                const int64_t Jp = (p - 1) * PT, rows_p = np - Jp * TILE, wp = (J - Jp) * TILE;
                const double* prev = cm->pbuf[(p - 1) & 1].p;
                const double* mine = prev + (J - Jp) * TILE;  // rows of this panel's diagonal block inside the previous buffer
                gemm(Ajj, m->cap, mine, rows_p, mine, rows_p, wc, wc, wp, -1.0, 1, 1, 0, pc);
                if (below > 0) {
                    CU(m, cudaEventRecord(m->evC, m->st2));
                    CU(m, cudaStreamWaitEvent(m->st3, m->evC, 0));
                    if (w.oz_digits && np - J * TILE >= OZ_MIN_ROWS) {
                        // the previous panel went through the tcgen05 update (same rule, same per-tile arithmetic as the single-GPU
                        // schedule, where these tiles belong to the main-stream launch): its digit slices start at this panel's rows
                        CU(m, cudaStreamWaitEvent(m->st3, m->evD, 0));
                        GemmArgs g{};
                        g.C = A21; g.ldc = m->cap;
                        g.M = (int)below; g.N = (int)wc; g.K = (int)wp;
                        g.alpha = -1.0; g.beta_one = 1; g.lower = 0;
                        const int KS = (int)(wp / OZ_KSTEP);
                        cnt.launches += ozaki_update_launch(g, oz_dig(p - 1) + (Jend - J) * (int64_t)KS * OZ_PART_BYTES, oz_sc(p - 1) + (Jend - J) * TILE,
                                                            oz_dig(p - 1), oz_sc(p - 1), 0, sc) > 0;
                    } else {
                        gemm(A21, m->cap, prev + (Jend - Jp) * TILE, rows_p, mine, rows_p, below, wc, wp, -1.0, 1, 0, 0, sc);
                    }
                    CU(m, cudaEventRecord(m->evC, m->st3));
                }
            }
            mark(p, 0, m->st2);
            launch_potrf_head(Ajj, m->cap, (int)(Jend - J), w.inv + J * TILE * TILE, w.W + p * HEAD_PANEL * HEAD_PANEL, w.P,
                              w.sync + p * HEAD_SYNC_INTS, has_eps, eps, m->info_d, (int)(J * TILE), pc);
            cnt.launches += 1;
            mark(p, 1, m->st2);
            CU(m, cudaMemcpy2DAsync(buf, rows * sizeof(double), Ajj, m->cap * sizeof(double), wc * sizeof(double), (size_t)wc,
                                    cudaMemcpyDeviceToDevice, m->st2));
            if (below > 0) {
                if (p >= 1) CU(m, cudaStreamWaitEvent(m->st2, m->evC, 0));
                gemm(buf + wc, rows, A21, m->cap, w.W + p * HEAD_PANEL * HEAD_PANEL, HEAD_PANEL, below, wc, wc, 1.0, 0, 0, 1, pc);
            }
            mark(p, 2, m->st2);
            CU(m, cudaEventRecord(cm->ev_col, m->st2));
            CU(m, cudaStreamWaitEvent(cm->st_comm, cm->ev_col, 0));
            if (below > 0) {  // L21 <- buffer, off the critical path (nothing reads these columns of L before the fit ends)
                CU(m, cudaStreamWaitEvent(m->st3, cm->ev_col, 0));
                CU(m, cudaMemcpy2DAsync(A21, m->cap * sizeof(double), buf + wc, rows * sizeof(double), below * sizeof(double),
                                        (size_t)wc, cudaMemcpyDeviceToDevice, m->st3));
                CU(m, cudaEventRecord(cm->ev_copy[p & 1], m->st3));
                copy_pending[p & 1] = true;
            }
        }
        if (P > 1) {
            // profiled (class "other") only on the owner: there the duration is the send itself, elsewhere it includes the wait
            // for the owner's factorisation; the "flops" slot carries the bytes
            LaunchCtx bc = mc;
            bc.st = cm->st_comm;
            if (owner != r) bc.prof = nullptr;
            ProfScope ps(bc, PROF_OTHER, (double)rows * wc * sizeof(double));
            NC(m, nccl->Broadcast(buf, buf, (size_t)rows * wc, ncclDouble, owner, cm->comm, cm->st_comm));
            cm->bcast_bytes += (double)rows * wc * sizeof(double);
            // the panel's inverse diagonal block W_p = L11^-1 (2 MB) follows: with it every rank can run the panel solves of
            // predict and of the LML gradient locally (trsm.cuh), like after a single-GPU fit
            double* Wp = w.W + p * HEAD_PANEL * HEAD_PANEL;
            NC(m, nccl->Broadcast(Wp, Wp, (size_t)HEAD_PANEL * HEAD_PANEL, ncclDouble, owner, cm->comm, cm->st_comm));
            cm->bcast_bytes += (double)HEAD_PANEL * HEAD_PANEL * sizeof(double);
        }
        mark(p, 3, cm->st_comm);
        CU(m, cudaEventRecord(cm->ev_bcast, cm->st_comm));
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_bcast, 0));  // the next owner's look-ahead reads the buffer on st2 / st3
        CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
        mark(p, 4, m->st);
        // tcgen05 updates (csrc/ozaki.cuh) while >= OZ_MIN_ROWS rows are left below the panel: every rank slices the rows below the
        // diagonal block into base-128 digits (tile 0 of the digit buffer = first row below the panel)
        const bool oz = w.oz_digits && below >= OZ_MIN_ROWS;
        if (oz) {
            if (w.oz_off_bytes) {
                // every panel has its own digit image, so the slicing does not have to queue behind the main stream's update with
                // the PREVIOUS panel: it runs on the side stream as soon as the panel has arrived — the next owner's look-ahead
                // (which reads these digits) then waits for the transfer only, not for this rank's trailing update
                CU(m, cudaStreamWaitEvent(m->st3, cm->ev_bcast, 0));
                ozaki_slice_launch(buf + wc, rows, below, (int)wc, oz_dig(p), oz_sc(p), sc);
                CU(m, cudaEventRecord(m->evD, m->st3));
                CU(m, cudaStreamWaitEvent(m->st, m->evD, 0));
            } else {
                ozaki_slice_launch(buf + wc, rows, below, (int)wc, oz_dig(p), oz_sc(p), mc);
                CU(m, cudaEventRecord(m->evD, m->st));
            }
            cnt.launches += 2;
        }
        {
            // every owned panel c > p, c != p+1, in ONE launch: the owned panels are groups of PT tile columns, P*PT apart
            int64_t c_first = p + 1 + ((r - (p + 1)) % P + P) % P;
            if (c_first == p + 1) c_first += P;
            if (c_first < NP) {
                const int64_t c0 = c_first * PT;
                int64_t ncols = 0;
                for (int64_t c = c_first; c < NP; c += P) ncols += std::min<int64_t>(PT, nb - c * PT);
                GemmArgs g{};
                g.C = m->L.p + c0 * TILE + c0 * TILE * m->cap; g.ldc = m->cap;
                g.A = buf + (c0 - J) * TILE; g.lda = rows;
                g.B = g.A; g.ldb = rows;
                g.M = (int)(np - c0 * TILE); g.N = (int)(ncols * TILE); g.K = (int)wc;
                g.alpha = -1.0; g.beta_one = 1; g.lower = 1;
                g.grp = (int)PT; g.stride = (int)(P * PT);
                if (oz) {
                    const int64_t toff = c0 - Jend;   // C's origin in tiles below the panel
                    const int8_t* dg = oz_dig(p) + toff * (wc / OZ_KSTEP) * (int64_t)OZ_PART_BYTES;
                    cnt.launches += ozaki_update_launch(g, dg, oz_sc(p) + toff * TILE, dg, oz_sc(p) + toff * TILE, P > 1 ? OZ_SHARDED_TPC : 0, mc) > 0;
                } else {
                    cnt.launches += gemm_nt_launch(g, mc) > 0;
                }
            }
        }
        if (owner != r)
            CU(m, cudaMemcpy2DAsync(m->L.p + J * TILE + J * TILE * m->cap, m->cap * sizeof(double), buf, rows * sizeof(double),
                                    rows * sizeof(double), (size_t)wc, cudaMemcpyDeviceToDevice, m->st));
        CU(m, cudaEventRecord(cm->ev_trail[p & 1], m->st));
        mark(p, 5, m->st);
    }
    if (trace && !tev.empty()) {
        cudaStreamSynchronize(m->st);
        cudaStreamSynchronize(m->st2);
        cudaStreamSynchronize(cm->st_comm);
        fprintf(stderr, "[shard trace rank %d] panel owner | lookahead-done(head start) head-done solve-done | bcast-done | main-start main-end (ms)\n", r);
        for (int64_t pnl = 0; pnl < NP; ++pnl) {
            float t[6] = {-1, -1, -1, -1, -1, -1};
            for (int k = 0; k < 6; ++k)
                if (cudaEventQuery(tev[(size_t)pnl * 6 + k]) == cudaSuccess) cudaEventElapsedTime(&t[k], tev.back(), tev[(size_t)pnl * 6 + k]);
            cudaGetLastError();
            fprintf(stderr, "[shard trace rank %d] %3lld %d | %8.3f %8.3f %8.3f | %8.3f | %8.3f %8.3f\n", r, (long long)pnl, shard_owner(pnl, P), t[0], t[1], t[2],
                    t[3], t[4], t[5]);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    // join the panel and side streams; the inverted diagonal tiles live with their owners and every rank needs them for the solves
    CU(m, cudaEventRecord(cm->ev_col, m->st2));
    CU(m, cudaStreamWaitEvent(m->st, cm->ev_col, 0));
    for (int i = 0; i < 2; ++i)
        if (copy_pending[i]) CU(m, cudaStreamWaitEvent(m->st, cm->ev_copy[i], 0));
    if (P > 1) {
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[(NP - 1) & 1], 0));
        NC(m, nccl->GroupStart());
        for (int64_t p = 0; p < NP; ++p) {
            const int64_t J = p * PT, Jend = std::min<int64_t>(J + PT, nb);
            NC(m, nccl->Broadcast(m->inv.p + J * TILE * TILE, m->inv.p + J * TILE * TILE, (size_t)(Jend - J) * TILE * TILE, ncclDouble,
                                shard_owner(p, P), cm->comm, m->st2));
        }
        NC(m, nccl->GroupEnd());
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 0);
        NC(m, nccl->AllReduce(m->info_d, m->info_d, 1, ncclInt, ncclMin, cm->comm, m->st2));  // first failing column anywhere
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 1);
        CU(m, cudaEventRecord(cm->ev_bcast, m->st2));
        CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
    }
    launch_transpose_tiles(m->inv.p, m->invT.p, nb, m->st);
    m->launches += cnt.launches + 1;
    return FGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The head schedule with the panel travelling in ROW PIECES (default at every rank count; FGP_SHARD_PIPE=0 selects
// factor_sharded_head above).  The sharded fit is bound by a serial chain per panel — the owner's full-height look-ahead
// update, head, panel solve, the broadcast, the digit slicing on every rank — whose size-proportional steps each wait for the
// WHOLE previous step there.  Here the rows below a panel's diagonal block are cut into pieces (512, 512, then <= PIPE_ROWS
// rows each), every piece contiguous in the panel buffer ([L11 | piece 0 | piece 1 | ...], each column-major with its own
// leading dimension), and each piece moves on its own through
//     solve (owner, panel stream) -> broadcast (comm stream) -> digit slicing (side stream, every rank)
//                                 -> look-ahead update of the next panel's rows in that piece (side stream, next owner)
// so that the four stages overlap piece by piece and across panels.  The next owner's head needs piece 0 only (the rows of
// its diagonal block): it starts as soon as L11 + piece 0 (4 MB) have arrived, while the big pieces of the panel are still
// being solved and shipped.  Per-tile arithmetic is unchanged (pieces partition the row tiles), so the factor stays
// bit-identical to the single-GPU one.  Panels with fewer than OZ_MIN_ROWS rows below them (the last two) travel in one piece
// in the old layout.  W_p follows the last piece (nothing in the fit waits for it).
int factor_sharded_pipe(fgp_model* m, const fgp_kernel_desc* kd, const KernelTraits& kt, double noise, int has_eps, double eps,
                        const PotrfWork& w) {
    fgp_comm* cm = m->comm;
    const int P = cm->nranks, r = cm->rank;
    const NcclApi* nccl = nccl_api();
    if (P > 1 && !nccl) return fail(m, FGP_ERR_COMM, "libnccl.so.2 could not be loaded");
    const int64_t np = m->np, nb = np / TILE, PT = HEAD_PANEL / TILE, NP = (nb + PT - 1) / PT;
    static const int64_t PIPE_ROWS = getenv("FGP_SHARD_PIPE_ROWS") ? std::max(1024, atoi(getenv("FGP_SHARD_PIPE_ROWS"))) / TILE * TILE : 8192;
    static const int OZ_SHARDED_TPC = getenv("FGP_SHARD_TPC") ? atoi(getenv("FGP_SHARD_TPC")) : 1;
    FGP_TRY(reserve_sharded(m));
    CU(m, cudaMemsetAsync(m->info_d, 0, sizeof(int), m->st));
    CU(m, cudaMemsetAsync(w.sync, 0, (size_t)NP * HEAD_SYNC_INTS * sizeof(int), m->st));
    cm->bcast_bytes = 0.0;
    for (int64_t p = r; p < NP; p += P) {  // Gram: only the block columns this rank owns (algebra/mod.rs:67-79 restricted to them)
        const int64_t J = p * PT, Jend = std::min<int64_t>(J + PT, nb);
        PairArgs pa{};
        pa.xa_c = pa.xb_c = m->xc.p;
        pa.xa_r = pa.xb_r = m->xr.p;
        pa.na = pa.nb = m->nc.p;
        pa.dp = (int)m->dp;
        pa.rows = pa.cols = np;
        pa.row_tile0 = (int)J;
        pa.col_tile0 = (int)(J * TILE / PAIR_TN);
        pa.col_tiles = (int)((Jend - J) * TILE / PAIR_TN);
        pa.symmetric = 1;
        write_covariance(m, kt, kd, pa, m->L.p, m->cap, m->n, m->n, noise * noise);
    }
    const LaunchCtx mc = m->ctx();
    LaunchCtx pc = mc, sc = mc;
    pc.st = m->st2;
    sc.st = m->st3;
    PotrfCounters cnt;
    CU(m, cudaEventRecord(m->evA, m->st));
    CU(m, cudaStreamWaitEvent(m->st2, m->evA, 0));
    CU(m, cudaStreamWaitEvent(m->st3, m->evA, 0));
    CU(m, cudaStreamWaitEvent(cm->st_comm, m->evA, 0));
    CU(m, cudaStreamWaitEvent(cm->st_copy, m->evA, 0));
    auto gemm = [&](double* Cp, int64_t ldc, const double* Ap, int64_t lda, const double* Bp, int64_t ldb, int64_t M, int64_t N,
                    int64_t K, double alpha, int beta_one, int lower, int k_upto, const LaunchCtx& c) {
        GemmArgs g{};
        g.C = Cp; g.ldc = ldc;
        g.A = Ap; g.lda = lda;
        g.B = Bp; g.ldb = ldb;
        g.M = (int)M; g.N = (int)N; g.K = (int)K;
        g.alpha = alpha; g.beta_one = beta_one; g.lower = lower; g.k_upto_col = k_upto;
        cnt.launches += gemm_nt_launch(g, c) > 0;
    };
    // a piece of the rows below a panel: rows [r0, r0 + h) counted from the first row below the diagonal block
    struct Piece { int64_t r0, h; double* ptr; int64_t ld; };
    struct Plan { int64_t J, Jend, rows, wc, below; bool piped, oz; double* buf; int64_t ld11; std::vector<Piece> pcs; };
    auto plan_of = [&](int64_t p) {
        Plan q;
        q.J = p * PT;
        q.Jend = std::min<int64_t>(q.J + PT, nb);
        q.rows = np - q.J * TILE;
        q.wc = (q.Jend - q.J) * TILE;
        q.below = q.rows - q.wc;
        q.buf = cm->pbuf[p & 1].p;
        q.oz = w.oz_digits && q.below >= OZ_MIN_ROWS;
        q.piped = q.oz && w.oz_off_bytes && w.oz_off_bytes[p] >= 0;
        q.ld11 = q.piped ? q.wc : q.rows;
        if (!q.piped) {
            if (q.below > 0) q.pcs.push_back({0, q.below, q.buf + q.wc, q.rows});
            return q;
        }
        double* at = q.buf + q.wc * q.wc;
        std::vector<std::pair<int64_t, int64_t>> rh;
        shard_pieces(q.below, PIPE_ROWS, rh);
        for (const auto& x : rh) {
            q.pcs.push_back({x.first, x.second, at, x.second});
            at += x.second * q.wc;
        }
        return q;
    };
    // events: [panel parity][kind][piece]
    enum { EV_SOLVE = 0, EV_BC = 1, EV_SL = 2, EV_LA = 3, EV_KINDS = 4 };
    int64_t max_pieces = 3 + (np / TILE * TILE + PIPE_ROWS - 1) / PIPE_ROWS + 1;
    if ((int64_t)cm->ev_pipe.size() < 2 * EV_KINDS * max_pieces) {
        const size_t old = cm->ev_pipe.size();
        cm->ev_pipe.resize((size_t)(2 * EV_KINDS * max_pieces), nullptr);
        for (size_t i = old; i < cm->ev_pipe.size(); ++i) CU(m, cudaEventCreateWithFlags(&cm->ev_pipe[i], cudaEventDisableTiming));
    }
    max_pieces = (int64_t)cm->ev_pipe.size() / (2 * EV_KINDS);
    auto ev = [&](int64_t p, int kind, size_t k) { return cm->ev_pipe[(size_t)(((p & 1) * EV_KINDS + kind) * max_pieces) + k]; };
    static const bool trace = getenv("FGP_SHARD_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](int64_t pnl, int what, cudaStream_t stm) {
        if (!trace) return;
        if (tev.empty()) {
            tev.resize((size_t)NP * 6 + 1);
            for (auto& e : tev) cudaEventCreate(&e);
            cudaEventRecord(tev.back(), m->st);
        }
        cudaEventRecord(tev[(size_t)pnl * 6 + what], stm);
    };
    auto oz_dig = [&](int64_t q) { return w.oz_digits + ((w.oz_off_bytes && w.oz_off_bytes[q] >= 0) ? w.oz_off_bytes[q] : 0); };
    auto oz_sc = [&](int64_t q) { return w.oz_scale + ((w.oz_off_rows && w.oz_off_rows[q] >= 0) ? w.oz_off_rows[q] : 0); };
    Plan prev{};
    for (int64_t p = 0; p < NP; ++p) {
        Plan cur = plan_of(p);
        const int64_t J = cur.J, Jend = cur.Jend, rows = cur.rows, wc = cur.wc, below = cur.below;
        const int owner = shard_owner(p, P);
        const bool next_mine = p + 1 < NP && shard_owner(p + 1, P) == r;
        double* buf = cur.buf;
        const size_t npc = cur.pcs.size();
        const int KS = (int)(wc / OZ_KSTEP);
        if (p >= 2) {
            // buffer p & 1 is free (panel p-2 has been sliced and copied into L) ...
            CU(m, cudaStreamWaitEvent(m->st2, cm->ev_copy[p & 1], 0));
            CU(m, cudaStreamWaitEvent(cm->st_comm, cm->ev_copy[p & 1], 0));
            // ... and this panel's columns have every update up to panel p-2: they are the FIRST group of this rank's update with
            // panel p-2, which is launched on its own so that the chain never waits for the bulk of a trailing update
            if (owner == r) CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[p & 1], 0));
        }
        if (owner == r) {
            double* Ajj = m->L.p + J * TILE * (m->cap + 1);
            double* A21 = m->L.p + Jend * TILE + J * TILE * m->cap;
            if (p >= 1) {
                // look-ahead, diagonal block: needs piece 0 of the previous panel only
                const Piece& p0 = prev.pcs[0];
                CU(m, cudaStreamWaitEvent(m->st2, ev(p - 1, EV_BC, 0), 0));
                gemm(Ajj, m->cap, p0.ptr, p0.ld, p0.ptr, p0.ld, wc, wc, prev.wc, -1.0, 1, 1, 0, pc);
                if (below > 0 && !prev.piped) {
                    // the previous panel came in one piece: rows below on the side stream beside the head (as factor_sharded_head)
                    CU(m, cudaEventRecord(m->evC, m->st2));
                    CU(m, cudaStreamWaitEvent(m->st3, m->evC, 0));
                    if (prev.oz) {
                        CU(m, cudaStreamWaitEvent(m->st3, m->evD, 0));
                        GemmArgs g{};
                        g.C = A21; g.ldc = m->cap;
                        g.M = (int)below; g.N = (int)wc; g.K = (int)prev.wc;
                        g.alpha = -1.0; g.beta_one = 1; g.lower = 0;
                        const int KSp = (int)(prev.wc / OZ_KSTEP);
                        cnt.launches += ozaki_update_launch(g, oz_dig(p - 1) + (Jend - J) * (int64_t)KSp * OZ_PART_BYTES, oz_sc(p - 1) + (Jend - J) * TILE,
                                                            oz_dig(p - 1), oz_sc(p - 1), 0, sc) > 0;
                    } else {
                        gemm(A21, m->cap, p0.ptr + wc, p0.ld, p0.ptr, p0.ld, below, wc, prev.wc, -1.0, 1, 0, 0, sc);
                    }
                    CU(m, cudaEventRecord(m->evC, m->st3));
                }
            }
            mark(p, 0, m->st2);
            launch_potrf_head(Ajj, m->cap, (int)(Jend - J), w.inv + J * TILE * TILE, w.W + p * HEAD_PANEL * HEAD_PANEL, w.P,
                              w.sync + p * HEAD_SYNC_INTS, has_eps, eps, m->info_d, (int)(J * TILE), pc);
            cnt.launches += 1;
            mark(p, 1, m->st2);
            CU(m, cudaMemcpy2DAsync(buf, cur.ld11 * sizeof(double), Ajj, m->cap * sizeof(double), wc * sizeof(double), (size_t)wc,
                                    cudaMemcpyDeviceToDevice, m->st2));
            if (npc == 0) CU(m, cudaEventRecord(ev(p, EV_SOLVE, 0), m->st2));
            if (p >= 1 && below > 0 && !prev.piped) CU(m, cudaStreamWaitEvent(m->st2, m->evC, 0));
            for (size_t k = 0; k < npc; ++k) {
                const Piece& pk = cur.pcs[k];
                if (p >= 1 && prev.piped)   // these rows have the previous panel's update: its pieces j >= 1 that overlap (prev row = row + wc)
                    for (size_t j = 1; j < prev.pcs.size(); ++j)
                        if (prev.pcs[j].r0 < pk.r0 + wc + pk.h && pk.r0 + wc < prev.pcs[j].r0 + prev.pcs[j].h)
                            CU(m, cudaStreamWaitEvent(m->st2, ev(p - 1, EV_LA, j), 0));
                gemm(pk.ptr, pk.ld, A21 + pk.r0, m->cap, w.W + p * HEAD_PANEL * HEAD_PANEL, HEAD_PANEL, pk.h, wc, wc, 1.0, 0, 0, 1, pc);
                CU(m, cudaEventRecord(ev(p, EV_SOLVE, k), m->st2));
            }
            mark(p, 2, m->st2);
        }
        // transfers: [L11 | piece 0] (or, in one piece, the whole buffer), then the other pieces, then W_p
        const size_t nreg = cur.piped ? npc : 1;
        for (size_t k = 0; k < nreg; ++k) {
            if (owner == r) CU(m, cudaStreamWaitEvent(cm->st_comm, ev(p, EV_SOLVE, cur.piped ? k : (npc ? npc - 1 : 0)), 0));
            if (P > 1) {
                double* src = (k == 0) ? buf : cur.pcs[k].ptr;
                const size_t count = !cur.piped ? (size_t)rows * wc : (k == 0 ? (size_t)(wc * wc + cur.pcs[0].h * wc) : (size_t)(cur.pcs[k].h * wc));
                LaunchCtx bc = mc;
                bc.st = cm->st_comm;
                if (owner != r) bc.prof = nullptr;   // profiled (class "other") on the owner only; the "flops" slot carries the bytes
                ProfScope ps(bc, PROF_OTHER, (double)count * sizeof(double));
                NC(m, nccl->Broadcast(src, src, count, ncclDouble, owner, cm->comm, cm->st_comm));
                cm->bcast_bytes += (double)count * sizeof(double);
            }
            CU(m, cudaEventRecord(ev(p, EV_BC, k), cm->st_comm));
        }
        mark(p, 3, cm->st_comm);
        if (P > 1) {
            // the panel's inverse diagonal block W_p = L11^-1 (2 MB): with it every rank runs the panel solves of predict and of the
            // LML gradient locally (trsm.cuh), like after a single-GPU fit
            double* Wp = w.W + p * HEAD_PANEL * HEAD_PANEL;
            if (owner == r) CU(m, cudaStreamWaitEvent(cm->st_comm, ev(p, EV_SOLVE, npc ? npc - 1 : 0), 0));
            NC(m, nccl->Broadcast(Wp, Wp, (size_t)HEAD_PANEL * HEAD_PANEL, ncclDouble, owner, cm->comm, cm->st_comm));
            cm->bcast_bytes += (double)HEAD_PANEL * HEAD_PANEL * sizeof(double);
        }
        // digit slices (every rank) and the next owner's look-ahead below its diagonal block, piece by piece on the side stream
        if (cur.piped) {
            if (next_mine && p >= 1) CU(m, cudaStreamWaitEvent(m->st3, cm->ev_trail[(p - 1) & 1], 0));  // the next panel's columns have panel p-1 (first group of that update)
            for (size_t k = 0; k < npc; ++k) {
                const Piece& pk = cur.pcs[k];
                CU(m, cudaStreamWaitEvent(m->st3, ev(p, EV_BC, k), 0));
                const int64_t t0 = pk.r0 / TILE;
                ozaki_slice_launch(pk.ptr, pk.ld, pk.h, (int)wc, oz_dig(p) + t0 * KS * (int64_t)OZ_PART_BYTES, oz_sc(p) + pk.r0, sc);
                cnt.launches += 2;
                CU(m, cudaEventRecord(ev(p, EV_SL, k), m->st3));
                if (next_mine && k >= 1) {
                    // rows of piece k in the next panel's columns -= (piece k)(piece 0)^T: the tiles the single-GPU schedule updates in
                    // its main-stream launch, same digits, same per-tile arithmetic
                    GemmArgs g{};
                    g.C = m->L.p + (Jend * TILE + pk.r0) + Jend * TILE * m->cap; g.ldc = m->cap;
                    g.M = (int)pk.h; g.N = (int)cur.pcs[0].h; g.K = (int)wc;
                    g.alpha = -1.0; g.beta_one = 1; g.lower = 0;
                    cnt.launches += ozaki_update_launch(g, oz_dig(p) + t0 * KS * (int64_t)OZ_PART_BYTES, oz_sc(p) + pk.r0, oz_dig(p), oz_sc(p), 0, sc) > 0;
                    CU(m, cudaEventRecord(ev(p, EV_LA, k), m->st3));
                }
            }
            CU(m, cudaStreamWaitEvent(m->st, ev(p, EV_SL, npc - 1), 0));
        } else {
            CU(m, cudaStreamWaitEvent(m->st, ev(p, EV_BC, 0), 0));
            if (cur.oz) {   // one scratch digit image: slice on the main stream, behind the update with the previous panel
                ozaki_slice_launch(buf + wc, rows, below, (int)wc, oz_dig(p), oz_sc(p), mc);
                CU(m, cudaEventRecord(m->evD, m->st));
                cnt.launches += 2;
            }
        }
        mark(p, 4, m->st);
        {
            // every owned panel c > p, c != p+1: the owned panels are groups of PT tile columns, P*PT apart.  The first group (the
            // panel this rank factors next) is its own launch followed by ev_trail: the look-ahead chain waits for that only.
            int64_t c_first = p + 1 + ((r - (p + 1)) % P + P) % P;
            if (c_first == p + 1) c_first += P;
            for (int part = 0; part < 2; ++part) {
                const int64_t cf = part == 0 ? c_first : c_first + P;
                if (cf < NP) {
                    const int64_t c0 = cf * PT;
                    int64_t ncols = 0;
                    for (int64_t c = cf; c < (part == 0 ? cf + 1 : NP); c += P) ncols += std::min<int64_t>(PT, nb - c * PT);
                    GemmArgs g{};
                    g.C = m->L.p + c0 * TILE + c0 * TILE * m->cap; g.ldc = m->cap;
                    g.A = buf + (c0 - J) * TILE; g.lda = rows;   // (read by the DMMA kernel only: one-piece layout)
                    g.B = g.A; g.ldb = rows;
                    g.M = (int)(np - c0 * TILE); g.N = (int)(ncols * TILE); g.K = (int)wc;
                    g.alpha = -1.0; g.beta_one = 1; g.lower = 1;
                    g.grp = (int)PT; g.stride = (int)(P * PT);
                    if (cur.oz) {
                        const int64_t toff = c0 - Jend;   // C's origin in tiles below the panel
                        const int8_t* dg = oz_dig(p) + toff * KS * (int64_t)OZ_PART_BYTES;
                        cnt.launches += ozaki_update_launch(g, dg, oz_sc(p) + toff * TILE, dg, oz_sc(p) + toff * TILE, P > 1 ? OZ_SHARDED_TPC : 0, mc) > 0;
                    } else {
                        cnt.launches += gemm_nt_launch(g, mc) > 0;
                    }
                }
                if (part == 0) CU(m, cudaEventRecord(cm->ev_trail[p & 1], m->st));
            }
        }
        // this rank's copy of the panel in L (owner: the solved rows below; L11 is in place), on the copy stream; then the buffer is
        // free for panel p+2 (one-piece panels without digit slices are read by the main-stream DMMA update as well)
        {
            cudaStream_t cs = cm->st_copy;
            CU(m, cudaStreamWaitEvent(cs, ev(p, EV_BC, nreg - 1), 0));
            if (owner != r)
                CU(m, cudaMemcpy2DAsync(m->L.p + J * TILE + J * TILE * m->cap, m->cap * sizeof(double), buf, cur.ld11 * sizeof(double),
                                        (cur.piped ? wc : rows) * sizeof(double), (size_t)wc, cudaMemcpyDeviceToDevice, cs));
            if (owner == r || cur.piped)
                for (size_t k = 0; k < npc; ++k)
                    CU(m, cudaMemcpy2DAsync(m->L.p + (Jend * TILE + cur.pcs[k].r0) + J * TILE * m->cap, m->cap * sizeof(double), cur.pcs[k].ptr,
                                            cur.pcs[k].ld * sizeof(double), cur.pcs[k].h * sizeof(double), (size_t)wc, cudaMemcpyDeviceToDevice, cs));
            if (cur.piped) {
                CU(m, cudaStreamWaitEvent(cs, ev(p, EV_SL, npc - 1), 0));
            } else {
                CU(m, cudaEventRecord(m->evB, m->st));   // the main stream's slicing / DMMA update have read the buffer
                CU(m, cudaStreamWaitEvent(cs, m->evB, 0));
                // (the next owner's look-ahead reads it in iteration p+1 on its panel / side stream, ahead of everything that writes
                // buffer p & 1 on that rank: its own transfers follow its solves)
            }
            CU(m, cudaEventRecord(cm->ev_copy[p & 1], cs));
        }
        mark(p, 5, m->st);
        prev = std::move(cur);
    }
    if (trace && !tev.empty()) {
        cudaStreamSynchronize(m->st);
        cudaStreamSynchronize(m->st2);
        cudaStreamSynchronize(m->st3);
        cudaStreamSynchronize(cm->st_comm);
        fprintf(stderr, "[shard trace rank %d] panel owner | lookahead-done(head start) head-done solve-done | bcast-done | main-start main-end (ms)\n", r);
        for (int64_t pnl = 0; pnl < NP; ++pnl) {
            float t[6] = {-1, -1, -1, -1, -1, -1};
            for (int k = 0; k < 6; ++k)
                if (cudaEventQuery(tev[(size_t)pnl * 6 + k]) == cudaSuccess) cudaEventElapsedTime(&t[k], tev.back(), tev[(size_t)pnl * 6 + k]);
            cudaGetLastError();
            fprintf(stderr, "[shard trace rank %d] %3lld %d | %8.3f %8.3f %8.3f | %8.3f | %8.3f %8.3f\n", r, (long long)pnl, shard_owner(pnl, P), t[0], t[1], t[2],
                    t[3], t[4], t[5]);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    // join the panel, side and comm streams; the inverted diagonal tiles live with their owners and every rank needs them for the solves
    CU(m, cudaEventRecord(cm->ev_col, m->st2));
    CU(m, cudaStreamWaitEvent(m->st, cm->ev_col, 0));
    CU(m, cudaEventRecord(cm->ev_col, m->st3));
    CU(m, cudaStreamWaitEvent(m->st, cm->ev_col, 0));
    CU(m, cudaEventRecord(cm->ev_bcast, cm->st_comm));
    CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
    CU(m, cudaEventRecord(cm->ev_col, cm->st_copy));
    CU(m, cudaStreamWaitEvent(m->st, cm->ev_col, 0));
    if (P > 1) {
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_trail[(NP - 1) & 1], 0));
        CU(m, cudaStreamWaitEvent(m->st2, cm->ev_bcast, 0));   // NCCL calls of one communicator stay in one order: after the W_p transfers
        NC(m, nccl->GroupStart());
        for (int64_t p = 0; p < NP; ++p) {
            const int64_t J = p * PT, Jend = std::min<int64_t>(J + PT, nb);
            NC(m, nccl->Broadcast(m->inv.p + J * TILE * TILE, m->inv.p + J * TILE * TILE, (size_t)(Jend - J) * TILE * TILE, ncclDouble,
                                shard_owner(p, P), cm->comm, m->st2));
        }
        NC(m, nccl->GroupEnd());
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 0);
        NC(m, nccl->AllReduce(m->info_d, m->info_d, 1, ncclInt, ncclMin, cm->comm, m->st2));  // first failing column anywhere
        info_encode_kernel<<<1, 1, 0, m->st2>>>(m->info_d, 1);
        CU(m, cudaEventRecord(cm->ev_bcast, m->st2));
        CU(m, cudaStreamWaitEvent(m->st, cm->ev_bcast, 0));
    }
    launch_transpose_tiles(m->inv.p, m->invT.p, nb, m->st);
    m->launches += cnt.launches + 1;
    return FGP_OK;
}

}  // namespace fgp
