"""One multi-GPU visit inside ONE process group (every torchrun start-up costs N x a minute of box time):
  1. n = 20000 (BASELINE.json configs[2]) device-resident, several STARNEIG_B200_* settings: device time (max over ranks),
     phase split of rank 0, scaling against a given 1-GPU time;
  2. the reference driver's acceptance checks (tools/invariants.py) on the result of the first setting, gathered on rank 0;
  3. n = STARNEIG_BIG_N (50000, configs[3]) on all GPUs with the same checks.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/visit8.py
       STARNEIG_SWEEP="K1=V1,K2=V2;K3=V3;..." (keys without the STARNEIG_B200_ prefix; an empty entry is the default; the
       pseudo-key P=k runs the setting on the first k GPUs of the box only, the other ranks idle meanwhile),
       STARNEIG_T1_MS (1-GPU device time for the scaling column), STARNEIG_BIG_N=0 skips step 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import starneig_b200 as sn
from starneig_b200 import dist as sdist
from tools import invariants

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
t1 = float(os.environ.get("STARNEIG_T1_MS", "5062"))
configs = os.environ.get("STARNEIG_SWEEP", "").split(";")
big_n = int(os.environ.get("STARNEIG_BIG_N", "50000"))


def say(*a):
    if rank == 0:
        print(*a, flush=True)


groups = {world: None}
for k in (2, 4):
    if k < world:
        groups[k] = dist.new_group(list(range(k)))


def run_size(n, cfg, reps, check):
    added = []
    P = world
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        if k == "P":
            P = int(v)
            continue
        os.environ["STARNEIG_B200_" + k] = v
        added.append("STARNEIG_B200_" + k)
    if rank < P:
        run_on(n, cfg, reps, check, P, groups[P])
    for k in added:
        os.environ.pop(k, None)
    torch.cuda.empty_cache()
    dist.barrier()


def run_on(n, cfg, reps, check, world, grp):
    # (`world`, `grp`: the ranks that take part)
    ld = (n + 15) // 16 * 16
    gen = torch.Generator(device="cuda").manual_seed(2019)
    A0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)         # the same on every rank
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    sn.set_profile_level(1)
    L = sdist.init(n, group=grp)                        # the column block may be part of the configuration
    cols = torch.from_numpy(L.global_cols()).to(dev)
    A0loc = A0[cols].contiguous()
    if rank != 0 or not check:
        del A0
    q0, qrows = L.q_row0, L.q_rows
    ldq = (max(qrows, 1) + 15) // 16 * 16
    A = torch.empty_like(A0loc)
    Q = torch.zeros((n, ldq), dtype=torch.float64, device="cuda")
    qd = torch.arange(q0, q0 + qrows, device=dev)
    best, st_best = None, None
    for it in range(reps):                              # with reps > 1 the first one is the warm-up
        A.copy_(A0loc)
        Q.zero_()
        Q[qd, qd - q0] = 1.0
        torch.cuda.synchronize()
        dist.barrier(group=grp)
        assert sdist.hessenberg_device(n, A, ld, Q, ldq, group=grp) == 0
        st = sn.get_stats()
        t = torch.tensor([st["device_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=grp)
        ms = float(t.item())
        if (it > 0 or reps == 1) and (best is None or ms < best):
            best, st_best = ms, st
    st = st_best
    say(f"[{cfg or 'default':44s}] gpus {world} n {n} pw {st['panel_width_used']} device_ms {best:8.1f} TFLOP/s {10 / 3 * n ** 3 / best / 1e9:6.2f} "
        + (f"x{t1 / best:5.2f} vs {t1:.0f} ms  " if n == 20000 else "")
        + f"col {st['panel_ms']:7.1f} trail {st['trail_ms']:6.1f} deferred {st['other_ms']:6.1f} gemv_ms {st['gemv_ms']:7.1f} "
        f"({st['gemv_timed_bytes'] / max(st['gemv_ms'], 1e-9) / 1e6:5.0f} GB/s rank 0) ph A/A'/R/R' {[round(x) for x in st['fused_phase_ms']]}")
    if check:
        del A0loc
        assert grp is None, "the parity gather runs on the whole process group"
        Ht, Qt = invariants.gather_to_rank0(A, Q, n, ld, lambda r: sdist.Layout(world, r, n), dist)
        del A, Q
        torch.cuda.synchronize()
        dist.barrier()
    sdist.finalize()
    sn.starneig_node_finalize()
    if check and rank == 0:
        t0 = time.time()
        inv = invariants.evaluate(A0, Ht, Qt, n)
        say(f"    parity n {n} on {world} GPUs: {inv}  ({time.time() - t0:.1f} s on rank 0's GPU)")
        assert inv["ok"], inv
        del A0, Ht, Qt


check_first = os.environ.get("STARNEIG_CHECK_FIRST", "1") != "0"
for idx, cfg in enumerate(configs):
    run_size(20000, cfg, 3, check=(idx == 0 and check_first and "P=" not in cfg))
if big_n > 0:
    run_size(big_n, "", 1, check=True)
say("OK")
dist.destroy_process_group()
