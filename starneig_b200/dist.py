"""One process per GPU: the torch.distributed plumbing around the C ABI's ``starneig_b200_dist_*`` entry points.

torch.distributed (NCCL on the GPU box, gloo in CPU tests) is used for exactly two things: all-gathering the
64-byte CUDA IPC handles of the ranks' exchange arenas, and barriers / max-over-ranks timing in the benchmark.
The data path itself never calls NCCL: the kernels exchange the per-column GEMV sums, the panel and the
per-panel top-row products through NVLink peer memory (starneig_b200/csrc/panel.cuh, engine.cuh).

Layout (mirrors ``ColMap`` in csrc/panel.cuh): A is 1-D block-cyclic by columns, global column
``c`` lives on rank ``(c // col_block) % world`` and a rank stores its columns contiguously in ascending global
order; Q is split into row slabs.
"""
import ctypes

import numpy as np

from . import api

HANDLE_BYTES = 64


class Layout:
    """Column / row ownership of one rank (host-only arithmetic through the C ABI)."""

    def __init__(self, world, rank, n):
        cb, lc, q0, qr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        ret = api.lib().starneig_b200_dist_layout(world, rank, n, ctypes.byref(cb), ctypes.byref(lc),
                                                   ctypes.byref(q0), ctypes.byref(qr))
        if ret != 0:
            raise ValueError(f"starneig_b200_dist_layout({world}, {rank}, {n}) returned {ret}")
        self.world, self.rank, self.n = world, rank, n
        self.col_block, self.local_cols, self.q_row0, self.q_rows = cb.value, lc.value, q0.value, qr.value

    def global_cols(self):
        """global index of every local column, ascending"""
        lc = np.arange(self.local_cols)
        return ((lc // self.col_block) * self.world + self.rank) * self.col_block + lc % self.col_block

    def owner(self, c):
        return (c // self.col_block) % self.world


def exchange_handles(local_handle: bytes, group=None):
    """all-gather of the ranks' 64-byte IPC handles -> one bytes object, rank order"""
    import torch
    import torch.distributed as dist
    assert len(local_handle) == HANDLE_BYTES
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(local_handle), dtype=torch.uint8).to(device)
    out = [torch.empty(HANDLE_BYTES, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def init(n_max, panel_width_max=-1, group=None):
    """Collective: create this rank's engine on the current CUDA device and connect the ranks' exchange arenas."""
    import torch.distributed as dist
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)
    buf = ctypes.create_string_buffer(HANDLE_BYTES)
    ret = api.lib().starneig_b200_dist_init(world, rank, n_max, panel_width_max, buf)
    if ret != 0:
        raise RuntimeError(f"starneig_b200_dist_init returned {ret}")
    if world > 1:
        handles = exchange_handles(buf.raw, group)
        ret = api.lib().starneig_b200_dist_connect(handles)
        if ret != 0:
            raise RuntimeError(f"starneig_b200_dist_connect returned {ret}")
        dist.barrier(group)
    return Layout(world, rank, n_max)


def _rendezvous(group=None):
    """Host-side rendezvous before a collective reduction: the device-side waits between the ranks have a time-out of a few
    seconds, so the ranks should enter the call together (allocation, garbage collection or I/O may delay one of them)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group)


def hessenberg_device(n, A_loc, ldA, Q_loc, ldQ, begin=0, end=None, panel_width=-1, group=None):
    """Collective. A_loc: this rank's columns (torch CUDA, column-major ldA x local_cols); Q_loc: its row slab."""
    end = n if end is None else end
    _rendezvous(group)
    return api.lib().starneig_b200_dist_hessenberg_device(n, begin, end, panel_width, A_loc.data_ptr(), ldA,
                                                          Q_loc.data_ptr(), ldQ)


def hessenberg_host(n, A, ldA, Q, ldQ, begin=0, end=None, panel_width=-1, group=None):
    """Collective. A, Q: column-major float64 numpy arrays holding the WHOLE matrices in memory shared by the ranks."""
    end = n if end is None else end
    _rendezvous(group)
    return api.lib().starneig_b200_dist_hessenberg_host(n, begin, end, panel_width, A.ctypes.data, ldA, Q.ctypes.data, ldQ)


def finalize():
    api.lib().starneig_b200_dist_finalize()
