"""Development check of the multi-rank engine (threads of one process). usage: multi_check.py P n [pw] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("STARNEIG_B200_VIRTUAL_RANKS", "8")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import starneig_b200 as sn
from oracle.oracle import Oracle
P = int(sys.argv[1]); n = int(sys.argv[2]); pw = int(sys.argv[3]) if len(sys.argv) > 3 else -1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ora = Oracle()
A0, Q0, ld = ora.fullpos(n, 2019)
sn.starneig_node_init(-1, P, sn.STARNEIG_NO_MESSAGES)
print("gpus", sn.starneig_node_get_gpus(), flush=True)
conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = pw
for it in range(reps):
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    t0 = time.time()
    ret = sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld)
    st = sn.get_stats()
    print(f"ret {ret} wall {time.time()-t0:.3f}s device_ms {st['device_ms']:.1f} panel {st['panel_ms']:.1f} trail {st['trail_ms']:.1f} "
          f"other {st['other_ms']:.1f} h2d {st['h2d_ms']:.1f} d2h {st['d2h_ms']:.1f} launches {st['kernel_launches']} ranks {st['ranks']} "
          f"GFLOP/s(dev) {10/3*n**3/st['device_ms']/1e6:.0f} fused_panels {st['fused_panels']} fused_ms {st['fused_kernel_ms']:.1f} gemv_ms {st['gemv_ms']:.1f} "
          f"gemv GB/s {st['gemv_timed_bytes']/max(st['gemv_ms'],1e-9)/1e6:.0f} phases A/A'/R/R' {[round(x,1) for x in st['fused_phase_ms']]}", flush=True)
if n <= 3000:
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw)
    u = 2.0 ** -52
    print("form", ora.hessenberg_form_violations(n, A, ld), "eh/u", np.abs(A[:n] - A2[:n]).max() / np.abs(A2[:n]).max() / u,
          "eq/u", np.abs(Q[:n] - Q2[:n]).max() / u, "res", ora.residual_u(n, Q, ld, A, ld, A0, ld), "orth", ora.orthogonality_u(n, Q, ld))
else:
    print("form", ora.hessenberg_form_violations(n, A, ld), "res", ora.residual_u(n, Q, ld, A, ld, A0, ld), "orth", ora.orthogonality_u(n, Q, ld))
sn.starneig_node_finalize()
