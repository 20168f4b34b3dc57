"""Rank process of tests/test_gpu_multi.py::test_processes_torchrun (launched by torch.distributed.run).

Every rank builds the same reference-driver input on the host, keeps only its shards (block-cyclic columns of A, a
row slab of Q) on its GPU, runs the collective reduction, and rank 0 gathers the shards and checks them against the
CPU oracle. A second pass goes through the host-pointer variant on a matrix in POSIX shared memory."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import starneig_b200 as sn                      # noqa: E402
from starneig_b200 import dist as sdist         # noqa: E402
from oracle.oracle import Oracle                # noqa: E402

U = 2.0 ** -52


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=600)
    ap.add_argument("--panel", dest="pw", type=int, default=-1)
    ap.add_argument("--devices", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ["LOCAL_RANK"]) % max(1, args.devices)
    torch.cuda.set_device(dev)
    # gloo for the plumbing when ranks share a device (NCCL refuses two ranks on one GPU)
    backend = "nccl" if args.devices >= world else "gloo"
    dist.init_process_group(backend, device_id=torch.device("cuda", dev) if backend == "nccl" else None)
    n, pw = args.n, args.pw
    ora = Oracle()
    A0, Q0, ld = ora.fullpos(n, 2019)

    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    L = sdist.init(n, pw)
    cols = L.global_cols()
    ldl = (n + 15) // 16 * 16
    ldq = (max(L.q_rows, 1) + 15) // 16 * 16
    # column-major shards as (cols, ld) row-major torch tensors
    A_loc = torch.zeros((max(L.local_cols, 1), ldl), dtype=torch.float64, device="cuda")
    A_loc[: L.local_cols, :n] = torch.from_numpy(np.ascontiguousarray(A0[:n, cols].T)).cuda()
    Q_loc = torch.zeros((n, ldq), dtype=torch.float64, device="cuda")
    Q_loc[:, : L.q_rows] = torch.from_numpy(np.ascontiguousarray(Q0[L.q_row0: L.q_row0 + L.q_rows, :n].T)).cuda()
    dist.barrier()
    ret = sdist.hessenberg_device(n, A_loc, ldl, Q_loc, ldq, panel_width=pw)
    assert ret == 0, ret
    st = sn.get_stats()
    assert st["ranks"] == world and st["kernel_launches"] > 0

    # gather on rank 0 through the host
    parts = [None] * world
    dist.all_gather_object(parts, (cols, A_loc[: L.local_cols, :n].cpu().numpy().T, L.q_row0, Q_loc[:, : L.q_rows].cpu().numpy().T))
    if rank == 0:
        H = np.zeros((n, n)); Q = np.zeros((n, n))
        for c, a, q0, q in parts:
            H[:, c] = a
            Q[q0: q0 + q.shape[0], :] = q
        A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
        assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
        eh = np.abs(H - A2[:n]).max() / np.abs(A2[:n]).max()
        eq = np.abs(Q - Q2[:n]).max()
        assert eh <= 200 * n * U and eq <= 200 * n * U, (eh, eq)
        assert np.array_equal(H == 0.0, A2[:n] == 0.0)
        print(f"device shards: |H-H_oracle|/max|H| = {eh:.2e}, |Q-Q_oracle| = {eq:.2e}")

    # host-pointer variant on a matrix shared by the rank processes
    name = f"/dev/shm/starneig_b200_test_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        buf = np.lib.format.open_memmap(name, mode="w+", dtype=np.float64, shape=(2, n, ld))
        buf[0] = A0.T
        buf[1] = Q0.T
        buf.flush()
    dist.barrier()
    buf = np.load(name, mmap_mode="r+")
    hA, hQ = buf[0].T, buf[1].T                      # column-major (ld x n) views of the shared pages
    ret = sdist.hessenberg_host(n, hA, ld, hQ, ld, panel_width=pw)
    assert ret == 0, ret
    dist.barrier()
    if rank == 0:
        A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
        ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw)
        eh = np.abs(hA[:n] - A2[:n]).max() / np.abs(A2[:n]).max()
        eq = np.abs(hQ[:n] - Q2[:n]).max()
        assert eh <= 200 * n * U and eq <= 200 * n * U, (eh, eq)
        print(f"shared host arrays: |H-H_oracle|/max|H| = {eh:.2e}, |Q-Q_oracle| = {eq:.2e}")
    dist.barrier()
    del buf, hA, hQ
    if rank == 0:
        os.unlink(name)
        print("DIST_WORKER_OK")
    sdist.finalize()
    sn.starneig_node_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
