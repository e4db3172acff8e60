"""CPU tests of the multi-GPU host logic: the block-cyclic plan exported by the library (host-only entry point) and the
world_size-2 gloo plumbing that ships the NCCL id between processes.  No GPU, no compute calls."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from friedrich_b200 import sharded


@pytest.mark.parametrize("n", [100, 512, 4096, 16384, 32768, 33000])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_partitions_every_panel_exactly_once(n, world):
    plans = [sharded.shard_plan(n, world, r) for r in range(world)]
    n_panels = plans[0]["n_panels"]
    tiles = -(-n // 128)
    pt = 4  # csrc/potrf.cuh HEAD_PANEL: 512-column panels on any number of GPUs (so sharded and single-GPU factors agree bit for bit)
    assert plans[0]["panel_cols"] == 128 * pt
    assert n_panels == -(-tiles // pt)
    owned = sorted(p for pl in plans for p in pl["owned"])
    assert owned == list(range(n_panels))
    assert sum(pl["n_owned"] for pl in plans) == n_panels
    assert abs(sum(pl["flop_share"] for pl in plans) - 1.0) < 1e-12
    if n >= 16384:  # block-cyclic keeps the trailing-update work balanced for the sizes the configs use
        assert max(pl["flop_share"] for pl in plans) < 1.0 / world + (0.07 if world <= 4 else 0.06)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = sharded.exchange_id(dist, rank, lambda: bytes(range(128)))
        plan = sharded.shard_plan(32768, world, rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, plan["owned"])
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([len(uid), uid[5], sum(len(g) for g in gathered),
                                                                    len(set(p for g in gathered for p in g))]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_id_exchange_and_plan_agreement(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got[0] == 128 and got[1] == 5      # both ranks hold rank 0's id bytes
        assert got[2] == 64 and got[3] == 64      # 64 panels at n=32768, each owned exactly once across the ranks


def test_sharded_entry_points_fail_loudly_without_gpu():
    import ctypes as C
    from friedrich_b200 import _native as N
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = N._h()
    assert N.lib().fgp_create(0, C.byref(h)) == N.FGP_ERR_CUDA  # no device => no handle => no sharded fit, no fallback


@pytest.mark.parametrize("tm,tn,grp,stride", [(1, 1, 1, 1), (7, 7, 1, 1), (9, 4, 1, 1), (128, 128, 1, 1), (20, 8, 4, 4),
                                              (64, 16, 4, 8), (61, 15, 4, 16), (250, 32, 4, 32), (33, 9, 4, 12),
                                              (256, 256, 4, 4), (1100, 40, 4, 32)])
def test_lower_mode_block_to_tile_map_is_exact(tm, tn, grp, stride):
    """Every thread block of a (sharded) trailing-update launch must land on a distinct tile on or below the diagonal of
    exactly the owned tile columns — checked on the host copy of the device decode."""
    import ctypes as C
    from friedrich_b200 import _native as N
    expect = []
    for jl in range(tn):
        tj = (jl // grp) * stride + jl % grp
        expect += [(ti, tj) for ti in range(tj, tm)]
    cap = len(expect) + 8
    ti, tj = (C.c_int * cap)(), (C.c_int * cap)()
    tiles = N.lib().fgp_dbg_lower_tiles(tm * 128, tn * 128, grp, stride, ti, tj, cap)
    assert tiles == len(expect)
    got = [(ti[b], tj[b]) for b in range(tiles)]
    assert sorted(got) == sorted(expect)  # every owned tile exactly once
    # band rasterisation: tile rows never go back to an earlier band of 16 tile rows (A blocks of a band stay in L2)
    R = 16
    while -(-tm // R) > 64:
        R *= 2
    bands = [t[0] // R for t in got]
    assert bands == sorted(bands)


@pytest.mark.parametrize("tm,tn,grp,stride,skip", [(7, 7, 1, 1, 4), (9, 4, 1, 1, 4), (128, 128, 1, 1, 4), (20, 8, 4, 4, 4),
                                                   (61, 15, 4, 16, 4), (250, 32, 4, 32, 4), (33, 33, 1, 1, 3), (5, 5, 1, 1, 5),
                                                   (256, 256, 1, 1, 4), (1100, 40, 4, 32, 2)])
def test_lower_mode_tile_map_with_skipped_top_rows(tm, tn, grp, stride, skip):
    """The trailing update behind a panel leaves out the next panel's diagonal block (updated by its own, earlier launch on
    the panel stream): the first `skip` tile rows are not visited, everything else exactly once."""
    import ctypes as C
    from friedrich_b200 import _native as N
    expect = []
    for jl in range(tn):
        tj = (jl // grp) * stride + jl % grp
        expect += [(ti, tj) for ti in range(max(tj, skip), tm)]
    cap = len(expect) + 8
    ti, tj = (C.c_int * cap)(), (C.c_int * cap)()
    tiles = N.lib().fgp_dbg_lower_tiles_skip(tm * 128, tn * 128, grp, stride, skip, ti, tj, cap)
    assert tiles == len(expect)
    got = [(ti[b], tj[b]) for b in range(tiles)]
    assert sorted(got) == sorted(expect)


@pytest.mark.parametrize("pipe_rows", [1024, 4096, 8192])
def test_sharded_panel_row_pieces_partition_the_rows(pipe_rows):
    """factor_sharded_pipe ships the rows below a panel's diagonal block in pieces (csrc/sharded.cuh shard_pieces): they must
    partition [0, below) in whole 128-row tiles, start with the next panel's diagonal-block rows (512) and the 512 rows of the
    panel after it, and never exceed pipe_rows (rounded up to a tile) — for every panel of every problem size."""
    import ctypes as C
    from friedrich_b200 import _native as N
    cap = 512
    r0, h = (C.c_int64 * cap)(), (C.c_int64 * cap)()
    for below in list(range(1024, 40 * 1024 + 1, 128)) + [65536 - 512, 131072 - 512]:
        k = N.lib().fgp_dbg_shard_pieces(below, pipe_rows, r0, h, cap)
        assert k >= 2, (below, k)
        pcs = [(r0[i], h[i]) for i in range(k)]
        assert pcs[0] == (0, 512) and pcs[1] == (512, 512)
        at = 0
        for first, height in pcs:
            assert first == at and height > 0 and height % 128 == 0
            at += height
        assert at == below
        assert all(height <= pipe_rows + 127 for _, height in pcs[2:])
        rest = [height for _, height in pcs[2:]]
        assert not rest or max(rest) - min(rest) <= 128   # equal pieces up to one tile
    assert N.lib().fgp_dbg_shard_pieces(512, pipe_rows, r0, h, cap) == -1      # one-piece panels are not cut
    assert N.lib().fgp_dbg_shard_pieces(1024 + 64, pipe_rows, r0, h, cap) == -1
