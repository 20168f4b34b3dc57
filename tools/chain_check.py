"""BASELINE.json configs[4] ("sep_sm_full_chain": the new GPU Hessenberg stage feeding the downstream Schur stage,
eigenvalue agreement with the all-CPU chain) at a chosen size. The reference's Schur stage (src/schur, StarPU) cannot be
built in this image, so LAPACK dhseqr from the same OpenBLAS stands in for starneig_SEP_SM_Schur on BOTH sides:
    GPU chain:  starneig_SEP_SM_Hessenberg (this library, host buffers)  ->  dhseqr  ->  eigenvalues
    CPU chain:  the reference's own Hessenberg sources (oracle/_ref; else the oracle port)  ->  dhseqr  ->  eigenvalues
Acceptance (BASELINE.json): every eigenvalue of the GPU chain has a partner of the CPU chain within 1e-10 * ||A||_F.
Matrix: the reference example's generator (examples/sep_sm_full_chain.c:65-75), entries uniform in [-1, 1].
usage: chain_check.py [n] [--cpu-only] [--save-cpu FILE | --load-cpu FILE]
       (default n = 4000; n = 10000 needs ~10 min of host time for the two dhseqr runs. --save-cpu stores the eigenvalues of
       the CPU chain (with --cpu-only: on a machine without a GPU), --load-cpu reads them back instead of re-running the CPU
       chain, so that the GPU box only pays for its own half: tests/golden/chain_n10000_cpu_eigs.npz was made that way.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle, Reference

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cpu_only = "--cpu-only" in sys.argv
save_cpu = sys.argv[sys.argv.index("--save-cpu") + 1] if "--save-cpu" in sys.argv else None
load_cpu = sys.argv[sys.argv.index("--load-cpu") + 1] if "--load-cpu" in sys.argv else None
ora = Oracle()
cores = os.cpu_count() or 1
ora.set_threads(cores)
ld = (n // 8 + 1) * 8                       # the example's leading dimension
rng = np.random.default_rng(2019)
A0 = np.zeros((ld, n), order="F")
A0[:n] = 2.0 * rng.random((n, n)) - 1.0
Q0 = np.zeros((ld, n), order="F")
Q0[np.arange(n), np.arange(n)] = 1.0
normA = np.linalg.norm(A0[:n])


def nearest_gap(ev_a, ev_b):
    """max over a of min over b |a - b|, in blocks (n x n complex distances do not fit for large n)"""
    worst = 0.0
    for i in range(0, len(ev_a), 512):
        d = np.abs(ev_a[i:i + 512, None] - ev_b[None, :]).min(axis=1)
        worst = max(worst, float(d.max()))
    return worst


if load_cpu:
    saved = np.load(load_cpu)
    assert int(saved["n"]) == n and abs(float(saved["normA"]) - float(normA)) <= 1e-12 * float(normA), "the saved CPU chain belongs to another matrix"
    ev_cpu = saved["eigenvalues"]
    print(f"n {n}  CPU chain: eigenvalues loaded from {load_cpu} ({str(saved['how'])})", flush=True)
else:
    t0 = time.time()
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    if Reference.available():
        ref = Reference(); ref.set_threads(cores); ref.set_workers(1)
        assert ref.hessenberg(n, A2, ld, Q2, ld) == 0
        kind = "reference sources (oracle/_ref)"
    else:
        assert ora.hessenberg_port(n, A2, ld, Q2, ld) == 0
        kind = "oracle port"
    t_cpu_h = time.time() - t0
    t0 = time.time()
    ev_cpu = ora.eigenvalues(n, A2, ld)
    t_cpu_s = time.time() - t0
    how = f"Hessenberg [{kind}] {t_cpu_h:.1f} s, dhseqr {t_cpu_s:.1f} s, {cores} cores"
    print(f"n {n} cores {cores}  CPU chain: {how}", flush=True)
    if save_cpu:
        np.savez_compressed(save_cpu, n=n, normA=normA, eigenvalues=ev_cpu, how=how)
if cpu_only:
    sys.exit(0)

import starneig_b200 as sn
A, Q = A0.copy(order="F"), Q0.copy(order="F")
sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
assert sn.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0          # untimed first call (context creation)
A, Q = A0.copy(order="F"), Q0.copy(order="F")
t0 = time.time()
assert sn.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
t_gpu_h = time.time() - t0
st = sn.get_stats()
sn.starneig_node_finalize()
t0 = time.time()
ev_gpu = ora.eigenvalues(n, A, ld)
t_gpu_s = time.time() - t0
gap = max(nearest_gap(ev_gpu, ev_cpu), nearest_gap(ev_cpu, ev_gpu))
u = 2.0 ** -52
print(f"GPU chain: starneig_SEP_SM_Hessenberg {t_gpu_h:.2f} s (device {st['device_ms']:.0f} ms), dhseqr {t_gpu_s:.1f} s")
print(f"form_violations {ora.hessenberg_form_violations(n, A, ld)} residual {ora.residual_u(n, Q, ld, A, ld, A0, ld):.1f} u "
      f"orthogonality {ora.orthogonality_u(n, Q, ld):.1f} u")
print(f"eigenvalue agreement: max nearest-partner distance {gap:.3e} = {gap / normA:.3e} * ||A||_F (bound 1e-10)  "
      f"trace {abs(ev_gpu.sum() - np.trace(A0[:n])) / normA:.2e} * ||A||_F")
assert gap <= 1e-10 * normA
print("OK")
