"""Quick hardware check of every opt-in engine switch (no torch import: numpy + the C ABI only): correctness at a small size
against the default kernels / the oracle, then one timed reduction per switch at a medium size.
usage: variant_smoke.py [n_check] [n_time]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starneig_b200 as sn
from oracle.oracle import Oracle

n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
ora = Oracle()
U = 2.0 ** -52
VARIANTS = ["", "FUSED_LL=1", "FUSED_LL=1,FUSED_R=1", "FUSED_EVEN_ROWS=1", "FUSED_LL=1,FUSED_R=1,FUSED_EVEN_ROWS=1", "GEMV_KC=2048",
            "GEMV_PREFETCH=32", "GEMV_RESIDENT_KB=20480", "OVERLAP=2", "GEMM_OPT=1", "GEMM_OPT=2", "GEMM_OPT=3",
            "FUSED_LL=1,FUSED_R=1,FUSED_EVEN_ROWS=1,GEMV_KC=2048,GEMV_PREFETCH=32,GEMV_RESIDENT_KB=20480"]
if os.environ.get("VARIANTS"):
    VARIANTS = os.environ["VARIANTS"].split(";")
BITWISE = {"FUSED_LL=1", "GEMV_KC=2048", "GEMV_PREFETCH=32", "GEMV_RESIDENT_KB=20480", "GEMM_OPT=1", "GEMM_OPT=2", "GEMM_OPT=3"}


def run(n, pw, A0, Q0, ld, cfg):
    added = []
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ["STARNEIG_B200_" + k] = v
        added.append("STARNEIG_B200_" + k)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    conf = sn.starneig_hessenberg_init_conf()
    conf.panel_width = pw
    ret = sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld)
    st = sn.get_stats()
    sn.starneig_node_finalize()
    for k in added:
        os.environ.pop(k, None)
    assert ret == 0
    return A, Q, st


t0 = time.time()
if n_check > 0:
  n, pw = n_check, 200
  A0, Q0, ld = ora.fullpos(n, 2019)
  A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
  ora.set_threads(os.cpu_count() or 1)
  ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw)
  Ad, Qd, _ = run(n, pw, A0, Q0, ld, "")
  for cfg in VARIANTS:
      try:
          A, Q, st = run(n, pw, A0, Q0, ld, cfg)
          eh = np.abs(A[:n] - A2[:n]).max() / np.abs(A2[:n]).max() / (n * U)
          eq = np.abs(Q[:n] - Q2[:n]).max() / (n * U)
          bit = np.array_equal(A, Ad) and np.array_equal(Q, Qd)
          ok = eh <= 200 and eq <= 200 and ora.hessenberg_form_violations(n, A, ld) == 0 and (bit or cfg not in BITWISE)
          print(f"check n={n} [{cfg or 'default':88s}] {'ok ' if ok else 'BAD'} |dH|/(n u max|H|) {eh:6.2f} |dQ|/(n u) {eq:6.2f} bitwise_equal_to_default {bit}", flush=True)
      except Exception as e:                      # noqa: BLE001
          print(f"check n={n} [{cfg}] EXCEPTION {e!r}", flush=True)
print(f"checks done after {time.time() - t0:.1f} s", flush=True)

n, pw = n_time, -1
A0, Q0, ld = ora.fullpos(n, 2019)
run(n, pw, A0, Q0, ld, "")                      # warm-up (allocations, module load)
for cfg in VARIANTS:
    _, _, st = run(n, pw, A0, Q0, ld, cfg)
    print(f"time  n={n} [{cfg or 'default':88s}] device_ms {st['device_ms']:8.2f} col {st['panel_ms']:8.2f} trail {st['trail_ms']:7.2f} "
          f"deferred {st['other_ms']:7.2f} gemv_ms {st['gemv_ms']:8.2f} ph {[round(x, 1) for x in st['fused_phase_ms']]}", flush=True)
print(f"all done after {time.time() - t0:.1f} s", flush=True)
