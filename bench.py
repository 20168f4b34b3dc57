#!/usr/bin/env python
"""bench.py -- FP64 Hessenberg GFLOP/s (10 n^3 / 3) of the B200-native path, n = 20000 on N GPUs of one box.

A "step" is one full reduction (with Q) of a random dense n x n FP64 matrix.
  value     device-resident arm: A and Q (N > 1: every rank's shards) already in HBM when the timed region starts
            (starneig_b200_hessenberg_device / starneig_b200_dist_hessenberg_device); time = CUDA events on the launching
            stream, max over ranks.
  e2e       the reference-facing call on pinned HOST buffers: H2D of A and Q, the reduction and D2H of H and Q are all
            inside the timed region. N = 1: starneig_SEP_SM_Hessenberg(n, A, ldA, Q, ldQ). N > 1: the same call in ONE
            process after starneig_node_init(cores, N, ...) (one host thread of the library per GPU), timed in a child
            process of rank 0; the one-process-per-GPU host arm (starneig_b200_dist_hessenberg_host) is reported beside it
            as `e2e_process_per_gpu` and is the fallback if the child fails.
  parity    outside the timed region, at every N: the reference test driver's acceptance checks (exact-zero Hessenberg
            form, |Q H Q^T - A|_F / |A|_F and |Q Q^T - I|_F / sqrt(n) in units of u; tools/invariants.py) on the WHOLE last
            timed result, evaluated on rank 0's GPU, asserted <= min(500 u, 10 n u); and bitwise equality of the host-buffer
            result with the device-resident one.
  roofline  the dominant kernel (k_panel_fused: one persistent launch per panel that streams the trailing matrix once per
            panel column): algorithmic GEMV bytes of a launch (sum over its columns of 8 * rows * cols) / mean launch
            duration from CUDA events recorded around every launch during the timed steps, against the measured HBM copy
            bandwidth in MEASURED_PEAKS.json; plus the device-side phase timers and the whole-path roofline.
  cpu_baseline  the reference's own CPU sources (oracle/_ref, built against a StarPU stand-in) -- or the oracle port when
            _ref is absent -- on a bounded n = 6000 sample, under two schedules of the reference's task graph (one worker
            thread per core with sequential BLAS, as the reference runs; or insertion order with threaded OpenBLAS): the
            faster one is the value, both are listed; and LAPACK dgehrd + dormhr (the reference driver's `lapack` solver)
            on the same sample.
`--impl reference` times that CPU implementation instead and prints the same JSON line with the same `config`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 Hessenberg GFLOP/s (10n^3/3) at n=20k"
UNIT = "GFLOP/s"
CPU_SAMPLE_N = 6000          # bounded CPU sample (~10-15 s per reduction on the box's cores)
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback
FP64_DMMA_TFLOPS = 37.0      # measured DMMA issue peak (profiles/r1_probe_peaks.log)
FP64_CUBLAS_TFLOPS = 35.7    # measured cublasDgemm 8192^3 (profiles/r1_probe_peaks.log)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures (profiles/), relative to
# the algorithmic GEMV bytes of that launch.
# k_panel_fused<0, 0> at the automatic width 192 (profiles/r2_final2_ncu_full_fused.txt, panel i=384): 589.9 GB DRAM vs 588.1 GB
# algorithmic; the extra 0.3 % is the level-2 side traffic of the panel (Y, V, VT of 192 columns stay in L2 for most of it;
# at width 312, round 1, it was 1.6 %: profiles/r1_s4_ncu_full_fused_and_dgemm.txt)
FUSED_TRAFFIC_RATIO = 1.003
# k_col_gemv alone (profiles/r1_ncu_full_baseline.txt): 3.1807 GB DRAM vs 3.1757 GB algorithmic
GEMV_TRAFFIC_RATIO = 1.0015


def flops(n):
    return 10.0 / 3.0 * float(n) ** 3


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax = float(parts[1]); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


class SharedHostMatrices:
    """A and Q (n x ld, C order: row c = column c of the matrix) in one POSIX shared-memory file mapped and
    page-locked by every rank process; falls back to private pinned arrays when /dev/shm is too small."""

    def __init__(self, n, ld, rank, dist):
        import numpy as np
        import torch
        self.rank, self.dist, self.path, self.registered = rank, dist, None, []
        nbytes = 2 * n * ld * 8
        ok = 0
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                ok = int(st.f_bavail * st.f_frsize > nbytes + (1 << 30))
            except OSError:
                ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.broadcast(flag, 0)
        if int(flag.item()) == 0:
            self.private = torch.empty((2, n, ld), dtype=torch.float64).pin_memory()
            self.A, self.Q = self.private[0].numpy(), self.private[1].numpy()
            return
        self.path = "/dev/shm/starneig_b200_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getuid())
        if rank == 0:
            if os.path.exists(self.path):
                os.unlink(self.path)
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        self.map = np.memmap(self.path, dtype=np.float64, mode="r+", shape=(2, n, ld))
        self.A, self.Q = self.map[0], self.map[1]
        rt = torch.cuda.cudart()
        for arr in (self.A, self.Q):
            err = rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
            if int(err) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared host matrices failed: {err}")
            self.registered.append(arr.ctypes.data)

    def close(self):
        import torch
        rt = torch.cuda.cudart()
        for ptr in self.registered:
            rt.cudaHostUnregister(ptr)
        self.registered = []
        self.dist.barrier()
        if self.path and self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)


def workload_name(n, gpus):
    """`config.workload` of both arms (ours and --impl reference)"""
    return f"Hessenberg reduction with Q, random dense FP64 n={n} (BASELINE.json configs[2] at {gpus} GPU{'s' if gpus > 1 else ''})"


def cpu_lapack_run(n, threads):
    """dgehrd + dormhr with threaded OpenBLAS: the reference test driver's own `lapack` solver
    (reference test/hessenberg/solvers.c:227-271), the first CPU baseline of BASELINE.md section 4. Returns seconds."""
    from oracle.oracle import Oracle
    ora = Oracle()
    ora.set_threads(threads)
    A, Q, ld = ora.fullpos(n, 2019)
    t0 = time.perf_counter()
    ret = ora.hessenberg_lapack(n, A, ld, Q, ld)
    dt = time.perf_counter() - t0
    assert ret == 0
    return dt


SCHEDULES = {
    "task-parallel": "task graph executed by one worker thread per core under its data dependencies (oracle/ref_shim/mini_starpu.c: "
                     "sequential-consistency rule, three priority levels), the reference's default tile size for that many workers, "
                     "sequential BLAS inside the codelets -- how the reference itself runs on CPU cores",
    "blas-parallel": "task graph executed in insertion order by one thread (one worker: large tiles), parallelism from threaded "
                     "OpenBLAS inside the codelets",
}


def cpu_reference_run(n, threads, schedule="task-parallel"):
    """One reduction with the reference's CPU implementation on a fullpos matrix; returns (seconds, kind).
    `schedule` selects how the reference's task graph is executed on the host cores (SCHEDULES)."""
    from oracle.oracle import Oracle, Reference
    ora = Oracle()
    A, Q, ld = ora.fullpos(n, 2019)
    if Reference.available():
        ref = Reference()
        if schedule == "task-parallel":
            ref.set_threads(1)
            ref.set_workers(threads)
            ref.set_executors(threads)
        else:
            ref.set_threads(threads)
            ref.set_workers(1)
            ref.set_executors(0)
        try:
            t0 = time.perf_counter()
            ret = ref.hessenberg(n, A, ld, Q, ld)
            dt = time.perf_counter() - t0
        finally:
            ref.set_executors(0)
        kind = "reference"
    else:
        ora.set_threads(threads)
        t0 = time.perf_counter()
        ret = ora.hessenberg_port(n, A, ld, Q, ld)
        dt = time.perf_counter() - t0
        kind = "port"
    assert ret == 0
    return dt, kind


def cpu_reference_child(n, schedule, count, timeout_s):
    """`count` reductions under `schedule` in a CHILD process with a time limit (the task-parallel executor of the stand-in is
    the one piece of the CPU arm that could deadlock; a child keeps that from taking the bench line with it).
    Returns the list of seconds, or None when the child failed or ran out of time."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-child", schedule, "--steps", str(count), "--cpu-n", str(n)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        if out.returncode != 0:
            print(f"[bench] CPU reference child ({schedule}) failed: {out.stderr[-400:]}", file=sys.stderr)
            return None
        return json.loads(out.stdout.strip().splitlines()[-1])["seconds"]
    except (subprocess.TimeoutExpired, ValueError, KeyError, IndexError) as e:
        print(f"[bench] CPU reference child ({schedule}) gave no result: {e!r}", file=sys.stderr)
        return None


def run_cpu_child(args):
    cores = min(os.cpu_count() or 1, 64)
    secs = [cpu_reference_run(args.cpu_n, cores, args.cpu_child)[0] for _ in range(args.steps)]
    print(json.dumps({"seconds": secs}))


def cpu_reference_probe(n, threads):
    """Both schedules of the reference's task graph once each; returns (faster schedule, {schedule: seconds}, kind).
    A schedule that failed is missing from the dict."""
    t_blas, kind = cpu_reference_run(n, threads, "blas-parallel")
    secs = {"blas-parallel": t_blas}
    if kind == "reference":         # (the port has one schedule: threaded BLAS)
        t = cpu_reference_child(n, "task-parallel", 1, max(180.0, 8.0 * t_blas))
        if t:
            secs["task-parallel"] = t[0]
    return min(secs, key=secs.get), secs, kind


def shared_config(n, gpus):
    """`config` of BOTH arms (ours and --impl reference): the same dict, so that the driver compares like with like.
    Everything arm-specific (panel width, switches, the CPU arm's bounded sample) lives in other keys of the line."""
    return {"workload": workload_name(n, gpus), "n": n}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 64)
    n = args.cpu_n
    # the reference's task graph under both schedules, once each (untimed: they are also the first two warm-up steps);
    # the timed steps use the faster one
    schedule, probe, kind = cpu_reference_probe(n, cores)
    extra_warmup = max(0, args.warmup - len(probe))
    times = None
    if schedule == "task-parallel":
        t = cpu_reference_child(n, schedule, extra_warmup + args.steps, (extra_warmup + args.steps) * 4.0 * probe[schedule] + 120.0)
        if t:
            times = t[extra_warmup:]
        else:
            schedule = "blas-parallel"
    if times is None:
        for _ in range(extra_warmup):
            cpu_reference_run(n, cores, schedule)
        times = [cpu_reference_run(n, cores, schedule)[0] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    value = flops(n) / (ms * 1e-3) / 1e9
    # the reference test driver's other CPU solver (LAPACK dgehrd + dormhr, threaded BLAS) on the same sample, once
    lapack_s = cpu_lapack_run(n, cores)
    sample = (f"each step is one full reduction with Q of a fullpos n={n} matrix (a bounded sample: the n={args.n} workload "
              f"would take ~{(args.n / n) ** 3 * ms / 6e4:.0f} min per step at this rate); GFLOP/s = 10 n^3 / 3 of the SAMPLE / its time; "
              + (f"reference src/hessenberg + src/common built from source against a StarPU stand-in; schedule `{schedule}` "
                 f"(the faster of the two probed): {SCHEDULES[schedule]}"
                 if kind == "reference" else "oracle port, threaded OpenBLAS"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.n, args.gpus),
        "sample_n": n,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "schedule": schedule,
                         "schedules_probed": {k: {"value": flops(n) / v / 1e9, "unit": UNIT, "seconds": v} for k, v in probe.items()},
                         "lapack": {"value": flops(n) / lapack_s / 1e9, "unit": UNIT, "cores": cores, "seconds": lapack_s,
                                    "what": f"LAPACK dgehrd + dormhr (the reference driver's `lapack` solver), fullpos n={n}, threaded OpenBLAS"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_threads_child(args):
    """N > 1: the reference-facing call itself -- ONE process, starneig_node_init(cores, N, ...) + starneig_SEP_SM_Hessenberg
    on pinned host arrays, one host thread of the library per GPU (reference src/include/starneig/node.h:178,
    sep_sm.h:89-92). Runs as a child process of rank 0 after the ranks of the process-per-GPU arms have gone, so that a
    failure here cannot take the bench line with it. Prints one JSON object."""
    import numpy as np
    import torch
    import starneig_b200 as sn
    from tools import invariants
    n, N = args.n, args.gpus
    ld = (n + 15) // 16 * 16
    torch.cuda.set_device(0)
    gen = torch.Generator(device="cuda").manual_seed(2019)
    dA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)
    hostA0 = dA0.cpu()
    pinned = torch.empty((2, n, ld), dtype=torch.float64).pin_memory()
    hA, hQ = pinned[0].numpy().T, pinned[1].numpy().T
    diag = np.arange(n)
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, N, sn.STARNEIG_NO_MESSAGES)
    sn.set_profile_level(1)
    times, st = [], None
    for it in range(1 + args.steps):
        pinned[0].copy_(hostA0)
        pinned[1].zero_()
        hQ[diag, diag] = 1.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert sn.starneig_SEP_SM_Hessenberg(n, hA, ld, hQ, ld) == 0
        dt = 1e3 * (time.perf_counter() - t0)
        st = sn.get_stats()
        if it > 0:
            times.append(dt)
    sn.starneig_node_finalize()
    parity = invariants.evaluate(dA0, pinned[0].cuda(), pinned[1].cuda(), n)
    ms = sum(times) / len(times)
    print(json.dumps({"ms_per_step": ms, "value": flops(n) / (ms * 1e-3) / 1e9, "h2d_bytes_per_step": int(st["h2d_bytes"]),
                      "d2h_bytes_per_step": int(st["d2h_bytes"]), "ranks": int(st["ranks"]), "panel_width": int(st["panel_width_used"]),
                      "staging_overlapped": int(st["staging_overlapped"]), "device_ms": st["device_ms"], "parity": parity}))


def reference_call_e2e(args, world):
    """rank 0, N > 1: times the one-process call through a child process (run_threads_child); returns (dict | None, error | None)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--e2e-threads-child", "--gpus", str(world), "--n", str(args.n),
           "--steps", str(max(1, min(args.steps, 5)))]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
                                                             "TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT, env=env)
    except subprocess.TimeoutExpired:
        return None, "timed out after 420 s"
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        return None, f"exit code {r.returncode}: {(r.stderr or r.stdout)[-300:]}"
    return json.loads(lines[-1]), None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import starneig_b200 as sn
    from starneig_b200 import dist as sdist
    n = args.n
    ld = (n + 15) // 16 * 16
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    sn.set_profile_level(2)

    # The same random matrix on every rank (same seed); a rank keeps only its shards in HBM: its block-cyclic
    # columns of A (full height) and its row slab of Q. world == 1: the whole matrices.
    gen = torch.Generator(device="cuda").manual_seed(2019)
    dA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)   # column-major (ld x n), entries in [0,1)
    if world > 1:
        L = sdist.init(n)
        cols = torch.from_numpy(L.global_cols()).cuda()
        dA0 = dA0[cols].contiguous()                    # (local_cols, ld): this rank's columns
        q0, qrows = L.q_row0, L.q_rows
    else:
        q0, qrows = 0, n
    ldq = (max(qrows, 1) + 15) // 16 * 16
    dA = torch.empty_like(dA0)
    dQ = torch.zeros((n, ldq), dtype=torch.float64, device="cuda")      # rows [q0, q0+qrows) of Q, all n columns
    qdiag = torch.arange(q0, q0 + qrows, device="cuda")

    def reset_device():
        dA.copy_(dA0)
        dQ.zero_()
        dQ[qdiag, qdiag - q0] = 1.0

    def reduce_device():
        if world > 1:
            return sdist.hessenberg_device(n, dA, ld, dQ, ldq)
        return sn.hessenberg_device(n, dA, ld, dQ, ldq)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- device-resident arm ----------------
    for _ in range(args.warmup):
        reset_device()
        barrier()
        assert reduce_device() == 0
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, gemv_ms, gemv_bytes, gemv_launches, launches, phase = [], 0.0, 0.0, 0, 0, [0.0, 0.0, 0.0]
    gemv_tbytes, gemv_tlaunches = 0.0, 0
    fused_panels, panels, gemm_flops = 0, 0, 0.0
    fused_ph = [0.0] * 5
    side_tail, overlap = 0.0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        reset_device()
        barrier()
        assert reduce_device() == 0
        st = sn.get_stats()
        dev_ms.append(st["device_ms"])
        gemv_ms += st["gemv_ms"]; gemv_bytes += st["gemv_bytes"]; gemv_launches += st["gemv_launches"]
        gemv_tbytes += st["gemv_timed_bytes"]; gemv_tlaunches += st["gemv_timed_launches"]
        launches += st["kernel_launches"]; gemm_flops += st["gemm_flops"]
        fused_panels += st["fused_panels"]; panels += st["panels"]
        side_tail += st["side_tail_ms"]; overlap = st["overlap"]
        fused_ph = [a + b for a, b in zip(fused_ph, list(st["fused_phase_ms"]) + [st["fused_kernel_ms"]])]
        phase = [phase[0] + st["panel_ms"], phase[1] + st["trail_ms"], phase[2] + st["other_ms"]]
    barrier()
    wall_ms_per_step = 1e3 * (time.perf_counter() - t_wall0) / args.steps
    clocks = sampler.stop()
    # CUDA events on the launching stream of every rank (first to last kernel of the call), max over ranks
    ms_per_step = max_over_ranks(sum(dev_ms) / len(dev_ms))
    value = flops(n) / (ms_per_step * 1e-3) / 1e9
    launches = int(sum_over_ranks(launches))

    # ---------------- parity of the LAST TIMED result (outside the timed region) ----------------
    # The reference test driver's acceptance checks (test/common/hooks.c:258-353,434-487, checks.c:180-208) on the whole
    # n x n result: exact-zero Hessenberg form, |Q H Q^T - A|_F / |A|_F and |Q Q^T - I|_F / sqrt(n) in units of u,
    # evaluated on rank 0's GPU (world > 1: the shards are gathered there first). Bound: min(500 u, 10 n u).
    from tools import invariants
    if world == 1:
        parity = invariants.evaluate(dA0, dA, dQ, n)
    else:
        Ht, Qt = invariants.gather_to_rank0(dA, dQ, n, ld, lambda r: sdist.Layout(world, r, n), dist)
        parity = None
        if rank == 0:
            gen = torch.Generator(device="cuda").manual_seed(2019)
            A0full = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)
            parity = invariants.evaluate(A0full, Ht, Qt, n)
            del A0full, Ht, Qt
        barrier()
    if rank == 0:
        parity["checked"] = "last timed result of the device-resident arm, all n columns"
        assert parity["ok"], f"parity check failed: {parity}"

    # ---------------- end-to-end arm: host buffers through the reference-facing call ----------------
    # world == 1: starneig_SEP_SM_Hessenberg on pinned host arrays. world > 1: every rank process holds the host
    # matrices (same content) in pinned memory and moves ONLY its shards to its GPU and back inside the timed
    # region (starneig_b200_dist_hessenberg_host), so the bytes over PCIe add up to A + Q once in each direction.
    shm = None
    if args.no_e2e:
        if world > 1:
            sdist.finalize()
        sn.starneig_node_finalize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_per_step, "config": shared_config(n, world), "e2e": None,
                              "parity": parity,
                              "note": "development run (--no-e2e): device-resident arm only, not a bench line",
                              "phases_ms_per_step": {"column_loops": phase[0] / args.steps, "trailing_updates": phase[1] / args.steps,
                                                     "deferred_busy": phase[2] / args.steps},
                              "fused_phases_ms_per_step": [x / args.steps for x in fused_ph],
                              "gemv_ms_per_step": gemv_ms / args.steps, "gpu_launches": launches, "clocks": clocks}))
        return
    if world == 1:
        gen = torch.Generator(device="cuda").manual_seed(2019)
        hostA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen).cpu()
        pinned = torch.empty((2, n, ld), dtype=torch.float64).pin_memory()
        hA, hQ = pinned[0].numpy().T, pinned[1].numpy().T         # column-major (ld x n) views
        eye_diag = np.arange(n)

        def reset_host():
            pinned[0].copy_(hostA0)
            pinned[1].zero_()
            hQ[eye_diag, eye_diag] = 1.0
    else:
        # ONE host copy of A and Q for all rank processes (POSIX shared memory, page-locked by every rank), as the
        # C ABI words it: "host arrays holding the whole matrices in memory shared by the ranks". Each rank resets,
        # uploads and writes back only its own shards.
        shm = SharedHostMatrices(n, ld, rank, dist)
        hA, hQ = shm.A.T, shm.Q.T
        hA_t, hQ_t = torch.from_numpy(shm.A), torch.from_numpy(shm.Q)       # (n, ld): row c = column c of the matrix
        cols_cpu = cols.cpu()

        def reset_host():
            hA_t[cols_cpu] = dA0.cpu()
            hQ_t[:, q0:q0 + qrows] = 0.0
            hQ_t.diagonal()[q0:q0 + qrows] = 1.0

    sn.set_profile_level(1)
    e2e_ms, h2d, d2h = [], 0, 0
    for it in range(1 + args.steps):
        reset_host()
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            assert sdist.hessenberg_host(n, hA, ld, hQ, ld) == 0
        else:
            assert sn.starneig_SEP_SM_Hessenberg(n, hA, ld, hQ, ld) == 0
        dt = 1e3 * (time.perf_counter() - t0)
        st = sn.get_stats()
        if it > 0:
            e2e_ms.append(dt); h2d = st["h2d_bytes"]; d2h = st["d2h_bytes"]
    e2e_ms_per_step = max_over_ranks(sum(e2e_ms) / len(e2e_ms))
    h2d, d2h = int(sum_over_ranks(h2d)), int(sum_over_ranks(d2h))
    e2e_value = flops(n) / (e2e_ms_per_step * 1e-3) / 1e9
    # the host-buffer call must return what the device-resident arm computed from the same input: compare this rank's
    # shards of the last e2e result with the last device-arm result (same kernels, same order of operations: bitwise)
    if world == 1:
        e2e_equal = bool(torch.equal(pinned[0].cuda(), dA)) and bool(torch.equal(pinned[1].cuda(), dQ))
    else:
        e2e_equal = (bool(torch.equal(hA_t[cols_cpu].cuda(), dA))
                     and bool(torch.equal(hQ_t[:, q0:q0 + qrows].contiguous().cuda(), dQ[:, :qrows].contiguous())))
    e2e_equal = max_over_ranks(0.0 if e2e_equal else 1.0) == 0.0
    assert e2e_equal, "the host-buffer (e2e) result differs from the device-resident result"
    if world > 1:
        sdist.finalize()
        shm.close()
    sn.starneig_node_finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return

    # ---------------- N > 1: the reference-facing call (one process drives all GPUs) ----------------
    e2e_calls = {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    e2e_extra = {}
    if world == 1:
        e2e_calls["path"] = "starneig_node_init(cores, 1, ...) + starneig_SEP_SM_Hessenberg(n, A, ldA, Q, ldQ) on pinned host arrays"
    if world > 1:
        e2e_calls["path"] = ("one process per GPU (torchrun): starneig_b200_dist_hessenberg_host on one shared, page-locked host copy "
                             "of A and Q, every rank stages its own shards")
        if not args.no_threads_e2e:
            del dA, dQ, dA0
            torch.cuda.empty_cache()
            child, err = reference_call_e2e(args, world)
            if child is not None and child["parity"]["ok"] and child["ranks"] == world:
                e2e_extra["e2e_process_per_gpu"] = e2e_calls
                e2e_calls = {"value": child["value"], "unit": UNIT, "ms_per_step": child["ms_per_step"],
                             "h2d_bytes_per_step": child["h2d_bytes_per_step"], "d2h_bytes_per_step": child["d2h_bytes_per_step"],
                             "path": f"ONE process: starneig_node_init(cores, {world}, ...) + starneig_SEP_SM_Hessenberg(n, A, ldA, Q, ldQ) on "
                                     "pinned host arrays, one host thread of the library per GPU (the reference's own calling convention)",
                             "steps": max(1, min(args.steps, 5)), "panel_width": child["panel_width"],
                             "staging_overlapped": child["staging_overlapped"], "parity": child["parity"]}
            else:
                e2e_extra["e2e_reference_call_error"] = err or f"child result rejected: {child}"

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_kind = measured_peaks()
    # level-3 flops of a reduction with Q: 14/3 n^3 (trailing 2, top rows 2/3, Q 2); 4 n^3 when Q = I is
    # accumulated backward after the last panel (Q: 4/3 n^3; engine.cuh, Rank::reduce) -- the roofline follows the algorithm run
    q_backward = bool(st.get("q_backward", 0))
    def t_roof(l3_coeff):
        return 1e3 * (8.0 * (n - 1) * n * (2 * n - 1) / 6.0 / (peak * 1e9) + l3_coeff * n ** 3 / (FP64_CUBLAS_TFLOPS * 1e12)) / world
    t_roof_ms = t_roof(4.0 if q_backward else 14.0 / 3.0)
    if fused_panels == panels and panels > 0:
        # One persistent kernel per panel (k_panel_fused): a launch streams the trailing matrix once per panel
        # column. Algorithmic bytes per launch = sum over its columns of 8 * rows * cols; launch duration = CUDA
        # events around every launch on the launching stream (stats panel_ms), so the level-2 side work and the
        # grid barriers inside the kernel count against the achieved figure.
        achieved = gemv_bytes / phase[0] / 1e6
        bytes_per_launch = gemv_bytes / panels
        gemv_phase_gbs = gemv_tbytes / gemv_ms / 1e6 if gemv_ms > 0 else None
        roofline = {
            "bound": "hbm", "kernel": "k_panel_fused", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": FUSED_TRAFFIC_RATIO * bytes_per_launch if FUSED_TRAFFIC_RATIO else None,
            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
            "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_step": panels // args.steps,
            "mean_launch_us": 1e3 * phase[0] / panels,
            "share_of_step": phase[0] / (ms_per_step * args.steps),
            # device-side timers of the level-2 phases inside the kernel (each incl. its grid barrier)
            "phases_ms_per_step": dict(zip(["finish_update", "w2_reduce", "reflector", "scalars_s", "kernel_total"],
                                           [x / args.steps for x in fused_ph])),
            # the GEMV phases alone (device-side %globaltimer around them, incl. the grid barrier that ends each)
            "gemv_phase": {"achieved": gemv_phase_gbs, "frac": gemv_phase_gbs / peak if gemv_phase_gbs else None,
                           "ms_per_step": gemv_ms / args.steps, "columns_per_step": gemv_launches // args.steps},
        }
    else:
        # three kernels per column: gemv_ms covers the event-timed k_col_gemv launches only (every 8th column):
        # divide THEIR bytes by THEIR time
        achieved = gemv_tbytes / gemv_ms / 1e6 if gemv_ms > 0 else None          # GB/s
        bytes_per_launch = gemv_bytes / max(1, gemv_launches)
        roofline = {
            "bound": "hbm", "kernel": "k_col_gemv", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if achieved else None,
            "traffic": GEMV_TRAFFIC_RATIO * bytes_per_launch if GEMV_TRAFFIC_RATIO else None,
            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
            "algorithmic_bytes_per_launch": bytes_per_launch, "launches_per_step": gemv_launches // args.steps,
            "timed_launches_per_step": gemv_tlaunches // args.steps,
            "mean_launch_us": 1e3 * gemv_ms / max(1, gemv_tlaunches),
            "share_of_step": (gemv_ms * gemv_bytes / max(1.0, gemv_tbytes)) / (ms_per_step * args.steps),
        }
    # whole-path roofline: T_roof(n, P) = (B / BW_hbm + (8/3 + 2) n^3 / F64_peak) / P
    roofline["t_roof_ms"] = t_roof_ms
    roofline["t_roof_ms_forward_q"] = t_roof(14.0 / 3.0)
    roofline["level3_flops_per_step"] = gemm_flops / args.steps
    roofline["fp64_peak_tflops"] = FP64_CUBLAS_TFLOPS
    roofline["fp64_peak_source"] = "cublasDgemm 8192^3 measured on this pool (profiles/r1_probe_peaks.log)"
    roofline["path_frac"] = roofline["t_roof_ms"] / ms_per_step

    # ---------------- CPU baseline (bounded sample) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = min(os.cpu_count() or 1, 64)
        schedule, probe, kind = cpu_reference_probe(args.cpu_n, cores)
        dt = probe[schedule]
        lapack_s = cpu_lapack_run(args.cpu_n, cores)
        cpu = {"value": flops(args.cpu_n) / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"one full reduction with Q of a fullpos n={args.cpu_n} matrix ({dt:.1f} s, schedule `{schedule}`: the faster "
                         f"of the two the reference's task graph was run under); n={n} would take ~{(n / args.cpu_n) ** 3 * dt / 60:.0f} min at this rate",
               "schedule": schedule,
               "schedules_probed": {k: {"value": flops(args.cpu_n) / v / 1e9, "unit": UNIT, "seconds": v} for k, v in probe.items()},
               "lapack": {"value": flops(args.cpu_n) / lapack_s / 1e9, "unit": UNIT, "cores": cores, "seconds": lapack_s,
                          "what": f"LAPACK dgehrd + dormhr (the reference driver's `lapack` solver), fullpos n={args.cpu_n}"}}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": shared_config(n, world),
        "engine": {"panel_width": int(st["panel_width_used"]), "ld": ld,
                   "q_accumulation": "backward after the last panel (Q = I on entry: 4/3 n^3 flops)" if q_backward
                   else "forward, panel by panel (2 n^3 flops)",
                   # engine switches taken from the environment (none: the defaults of DESIGN.md section 4)
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("STARNEIG_B200_")},
                   "l2": "inputs (A, Q: 2 x %.1f GB) are larger than the 126 MB L2; no explicit flush" % (n * ld * 8 / 1e9),
                   "parallelism": "1 GPU" if world == 1 else
                   f"{world} GPUs, one process each: A 1-D block-cyclic by columns (block 64), Q by row slabs; per-column GEMV "
                   "sums and per-panel products exchanged by the kernels over NVLink peer memory (no NCCL on the data path)"},
        "parity": dict(parity, e2e_equals_device_bitwise=e2e_equal),
        "wall_ms_per_step": wall_ms_per_step,
        # column loops + trailing updates are the critical path; with overlap the deferred Q / top-row updates run
        # on a side stream concurrently with the next column loops ("deferred_busy" is that stream's busy time) and
        # only "deferred_tail" (end of the last trailing update -> end of the call) adds to the step
        "phases_ms_per_step": {"column_loops": phase[0] / args.steps, "trailing_updates": phase[1] / args.steps,
                               "deferred_busy": phase[2] / args.steps, "deferred_tail": side_tail / args.steps,
                               "overlap": overlap},
        "e2e": e2e_calls,
        **e2e_extra,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("STARNEIG_BENCH_N", "20000")))
    ap.add_argument("--cpu-n", type=int, default=CPU_SAMPLE_N)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-threads-e2e", action="store_true", help="N > 1: skip the one-process (thread per GPU) timing of the reference-facing call")
    ap.add_argument("--e2e-threads-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-child", default=None, choices=list(SCHEDULES), help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the host-buffer arm (development runs at sizes whose host copies do not fit comfortably)")
    args = ap.parse_args()
    if args.e2e_threads_child:
        run_threads_child(args)
    elif args.cpu_child:
        run_cpu_child(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
