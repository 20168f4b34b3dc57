"""One reduction through the one-process call on P ranks (thread per rank; STARNEIG_B200_VIRTUAL_RANKS lets ranks share devices)
with the driver's invariants -- a development aid for the multi-GPU path.   usage: multi_once.py n panel_width P"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
os.environ.setdefault("STARNEIG_B200_VIRTUAL_RANKS", "8")
import starneig_b200 as sn
from oracle.oracle import Oracle
ora = Oracle()
n, pw, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A0, Q0, ld = ora.fullpos(n, 2019)
A, Q = A0.copy(order="F"), Q0.copy(order="F")
sn.starneig_node_init(sn.STARNEIG_USE_ALL, P, sn.STARNEIG_NO_MESSAGES)
conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = pw
print("ret", sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld), flush=True)
print("residual_u", ora.residual_u(n, Q, ld, A, ld, A0, ld), "orthogonality_u", ora.orthogonality_u(n, Q, ld), "q_backward", sn.get_stats()["q_backward"], flush=True)
sn.starneig_node_finalize()
