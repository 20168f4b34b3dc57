"""CPU tests (no GPU) of the host-side logic of the multi-GPU path: the block-cyclic column / row-slab layout
exported by the C ABI (starneig_b200_dist_layout, ColMap in csrc/panel.cuh) and the torch.distributed plumbing
(world_size-2 gloo) that all-gathers the ranks' IPC handles."""
import os
import socket

import numpy as np
import pytest


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [1, 7, 64, 65, 1000, 20000])
def test_layout_partitions_columns_and_rows(sn, world, n):
    from starneig_b200.dist import Layout
    seen_cols = np.zeros(n, dtype=int)
    seen_rows = np.zeros(n, dtype=int)
    for r in range(world):
        L = Layout(world, r, n)
        cols = L.global_cols()
        assert len(cols) == L.local_cols
        assert np.all(np.diff(cols) > 0)                      # ascending => any global range is a contiguous local range
        assert np.all(cols < n)
        assert all(L.owner(int(c)) == r for c in cols[:: max(1, len(cols) // 50)])
        for lc in (0, L.local_cols // 2, L.local_cols - 1):
            if 0 <= lc < L.local_cols:
                assert sn.lib().starneig_b200_dist_global_col(world, r, L.col_block, int(lc)) == cols[lc]
        seen_cols[cols] += 1
        seen_rows[L.q_row0: L.q_row0 + L.q_rows] += 1
    assert np.all(seen_cols == 1) and np.all(seen_rows == 1)


def test_layout_rejects_bad_arguments(sn):
    assert sn.lib().starneig_b200_dist_layout(0, 0, 10, None, None, None, None) == sn.STARNEIG_INVALID_ARGUMENTS
    assert sn.lib().starneig_b200_dist_layout(9, 0, 10, None, None, None, None) == sn.STARNEIG_INVALID_ARGUMENTS
    assert sn.lib().starneig_b200_dist_layout(2, 2, 10, None, None, None, None) == sn.STARNEIG_INVALID_ARGUMENTS
    assert sn.lib().starneig_b200_dist_layout(2, 0, 0, None, None, None, None) == sn.STARNEIG_INVALID_ARGUMENTS


def test_dist_calls_need_init(sn):
    """the collective entry points refuse to run before starneig_b200_dist_init (no GPU is touched)"""
    assert sn.lib().starneig_b200_dist_connect(None) == sn.STARNEIG_NOT_INITIALIZED
    assert sn.lib().starneig_b200_dist_hessenberg_device(8, 0, 8, -1, None, 8, None, 8) == sn.STARNEIG_NOT_INITIALIZED
    assert sn.lib().starneig_b200_dist_hessenberg_host(8, 0, 8, -1, None, 8, None, 8) == sn.STARNEIG_NOT_INITIALIZED
    buf = bytes(64)
    assert sn.lib().starneig_b200_dist_init(2, 0, 100, -1, buf) == sn.STARNEIG_NOT_INITIALIZED   # node not initialised


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from starneig_b200.dist import exchange_handles, Layout, HANDLE_BYTES
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * HANDLE_BYTES
    allh = exchange_handles(mine)
    ok = len(allh) == world * HANDLE_BYTES and all(allh[HANDLE_BYTES * r: HANDLE_BYTES * (r + 1)] == bytes([r + 1]) * HANDLE_BYTES
                                                  for r in range(world))
    # every rank derives the same global picture from its own rank id only
    L = Layout(world, rank, 1000)
    counts = [None] * world
    dist.all_gather_object(counts, (L.local_cols, L.q_rows))
    ok = ok and sum(c for c, _ in counts) == 1000 and sum(q for _, q in counts) == 1000
    out[rank] = ok
    dist.destroy_process_group()


def test_handle_exchange_gloo_world2():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
