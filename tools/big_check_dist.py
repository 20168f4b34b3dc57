"""BASELINE.json configs[3]: n = 50000 (20 GB A + 20 GB Q) sharded over the GPUs of one box, one process per GPU, device
resident (starneig_b200_dist_hessenberg_device). After the reduction the shards are gathered on rank 0's GPU (NCCL
send/recv, outside the timed call) and the reference driver's invariants are evaluated there with torch FP64 matmuls --
the CPU oracle cannot reach this size: exact-zero Hessenberg form, residual |Q H Q^T - A|_F / |A|_F, orthogonality
|Q Q^T - I|_F / sqrt(n), trace.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/big_check_dist.py
       (size: STARNEIG_BENCH_N, default 50000; STARNEIG_CHECK=0 skips the gather + invariants and only times)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import starneig_b200 as sn
from starneig_b200 import dist as sdist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("STARNEIG_BENCH_N", "50000"))
check = os.environ.get("STARNEIG_CHECK", "1") != "0"
reps = int(os.environ.get("STARNEIG_REPS", "1"))
ld = (n + 15) // 16 * 16
u = 2.0 ** -52
dev = torch.device("cuda", local)

sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
sn.set_profile_level(1)
gen = torch.Generator(device="cuda").manual_seed(2019)
A0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)      # row c = column c; the same on every rank
L = sdist.init(n)
cols = torch.from_numpy(L.global_cols()).to(dev)
A0loc = A0[cols].contiguous()
if rank != 0 or not check:
    del A0
q0, qrows = L.q_row0, L.q_rows
ldq = (max(qrows, 1) + 15) // 16 * 16
A = torch.empty_like(A0loc)
Q = torch.zeros((n, ldq), dtype=torch.float64, device="cuda")
qd = torch.arange(q0, q0 + qrows, device=dev)

best = None
for it in range(reps):
    A.copy_(A0loc)
    Q.zero_()
    Q[qd, qd - q0] = 1.0
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.time()
    ret = sdist.hessenberg_device(n, A, ld, Q, ldq)
    wall = time.time() - t0
    st = sn.get_stats()
    assert ret == 0
    t = torch.tensor([st["device_ms"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(f"n {n} gpus {world} panel_width {st['panel_width']} panels {st['panels']} wall {wall:.1f}s device_ms(max over ranks) {ms:.0f} "
              f"GFLOP/s {10 / 3 * n ** 3 / ms / 1e6:.0f} col {st['panel_ms']:.0f} trail {st['trail_ms']:.0f} deferred {st['other_ms']:.0f} "
              f"gemv_ms {st['gemv_ms']:.0f} ({st['gemv_timed_bytes'] / max(st['gemv_ms'], 1e-9) / 1e6:.0f} GB/s on rank 0) "
              f"phases A/A'/R/R' {[round(x) for x in st['fused_phase_ms']]} launches(rank 0) {st['kernel_launches']}", flush=True)
    best = ms if best is None else min(best, ms)
del A0loc

if check:
    # ---- gather H (columns) and Q (row slabs) on rank 0
    if rank == 0:
        H = torch.empty((n, ld), dtype=torch.float64, device="cuda")
        Qf = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
        H[cols] = A
        Qf[:, q0:q0 + qrows] = Q[:, :qrows]
        for r in range(1, world):
            Lr = sdist.Layout(world, r, n)
            buf = torch.empty((Lr.local_cols, ld), dtype=torch.float64, device="cuda")
            dist.recv(buf, src=r)
            H[torch.from_numpy(Lr.global_cols()).to(dev)] = buf
            del buf
            ldq_r = (max(Lr.q_rows, 1) + 15) // 16 * 16
            buf = torch.empty((n, ldq_r), dtype=torch.float64, device="cuda")
            dist.recv(buf, src=r)
            Qf[:, Lr.q_row0:Lr.q_row0 + Lr.q_rows] = buf[:, :Lr.q_rows]
            del buf
    else:
        dist.send(A, dst=0)
        dist.send(Q, dst=0)
    del A, Q
    torch.cuda.synchronize()
    dist.barrier()

sdist.finalize()
sn.starneig_node_finalize()

if check and rank == 0:
    # matrices as torch sees them: M_t[c, r] = M(r, c), i.e. the transpose. H^T must be lower Hessenberg in torch's view.
    Ht, Qt, A0t = H[:, :n], Qf[:, :n], A0[:, :n]
    bad = 0
    for c0 in range(0, n, 4096):
        blk = Ht[c0:c0 + 4096]                 # columns c0.. of H; entries (r, c) with r > c + 1 must be exactly zero
        rows = torch.arange(n, device=dev)[None, :]
        cc = torch.arange(c0, min(n, c0 + 4096), device=dev)[:, None]
        bad += int(((rows > cc + 1) & (blk != 0.0)).sum())
    finite = bool(torch.isfinite(Ht).all()) and bool(torch.isfinite(Qt).all())
    tr = abs(float(Ht.diagonal().sum() - A0t.diagonal().sum())) / abs(float(A0t.diagonal().sum()))
    normA = float(torch.linalg.norm(A0t))
    T = Qt.T @ Ht                               # (Q H Q^T)^T = Q_t^T H_t Q_t in torch's view
    R = T @ Qt
    del T
    R -= A0t
    res = float(torch.linalg.norm(R)) / normA / u
    del R
    G = Qt.T @ Qt                               # (Q Q^T)^T
    G.diagonal().sub_(1.0)
    orth = float(torch.linalg.norm(G)) / n ** 0.5 / u
    print(f"finite {finite} form_violations {bad} residual {res:.1f} u orthogonality {orth:.1f} u trace_rel_err {tr:.2e} "
          f"(bounds: 10 n u = {10 * n} u, reference driver warn 500 u / fail 10000 u)", flush=True)
    assert finite and bad == 0 and res < 500 and orth < 500
    print("OK", flush=True)
dist.barrier()
dist.destroy_process_group()
