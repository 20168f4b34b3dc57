"""Device-resident reductions on N GPUs (one process per GPU, torchrun) under several STARNEIG_B200_* settings inside ONE
process group: the settings are read when a rank's engine is created (starneig_b200_dist_init), so every configuration
re-creates the engine but not the processes, the CUDA contexts or the NCCL communicator. One line per configuration with
the device time (max over ranks), the phase split of rank 0 and the scaling against a given 1-GPU time.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_sweep.py
       configurations: STARNEIG_SWEEP="K1=V1,K2=V2;K3=V3;..." (keys without the STARNEIG_B200_ prefix; an empty entry is the default)
       size: STARNEIG_BENCH_N (20000); STARNEIG_T1_MS: 1-GPU device time for the scaling column (5407)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import starneig_b200 as sn
from starneig_b200 import dist as sdist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("STARNEIG_BENCH_N", "20000"))
t1 = float(os.environ.get("STARNEIG_T1_MS", "5407"))
configs = os.environ.get("STARNEIG_SWEEP", "").split(";")
ld = (n + 15) // 16 * 16
dev = torch.device("cuda", local)
gen = torch.Generator(device="cuda").manual_seed(2019)
A0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)

for cfg in configs:
    added = []
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ["STARNEIG_B200_" + k] = v
        added.append("STARNEIG_B200_" + k)
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    sn.set_profile_level(1)
    L = sdist.init(n)                                   # the column block may be part of the configuration
    cols = torch.from_numpy(L.global_cols()).to(dev)
    A0loc = A0[cols].contiguous()
    q0, qrows = L.q_row0, L.q_rows
    ldq = (max(qrows, 1) + 15) // 16 * 16
    A = torch.empty_like(A0loc)
    Q = torch.zeros((n, ldq), dtype=torch.float64, device="cuda")
    qd = torch.arange(q0, q0 + qrows, device=dev)
    best, st_best = None, None
    for it in range(3):                                 # first one is the warm-up
        A.copy_(A0loc)
        Q.zero_()
        Q[qd, qd - q0] = 1.0
        torch.cuda.synchronize()
        dist.barrier()
        ret = sdist.hessenberg_device(n, A, ld, Q, ldq)
        assert ret == 0
        st = sn.get_stats()
        t = torch.tensor([st["device_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if it > 0 and (best is None or ms < best):
            best, st_best = ms, st
    gc0 = int(cols[0])
    ok = bool(torch.isfinite(A[: min(A.shape[0], 1024)]).all()) and float(A[0, gc0 + 2: n].abs().max()) == 0.0
    if rank == 0:
        st = st_best
        print(f"[{cfg or 'default':60s}] gpus {world} n {n} device_ms {best:8.1f} TFLOP/s {10 / 3 * n ** 3 / best / 1e9:6.2f} x{t1 / best:5.2f} vs {t1:.0f} ms  "
              f"col {st['panel_ms']:7.1f} trail {st['trail_ms']:6.1f} deferred {st['other_ms']:6.1f} gemv_ms {st['gemv_ms']:7.1f} "
              f"ph {[round(x) for x in st['fused_phase_ms']]} form_ok {ok}", flush=True)
    sdist.finalize()
    sn.starneig_node_finalize()
    for k in added:
        os.environ.pop(k, None)
    del A, Q, A0loc
    dist.barrier()
dist.destroy_process_group()
