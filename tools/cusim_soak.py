"""Randomised soak of the kernel-logic emulator build (tests/cusim): random sizes, panel widths, rank counts, emulated SM
counts, thread schedules (shuffled / lagging blocks) and random combinations of the opt-in engine switches, each reduction
in its own process, checked against the CPU oracle (entrywise 200 n u, exact-zero Hessenberg form, residual and
orthogonality). usage: cusim_soak.py [seed] [seconds]   (build first: make -C tests/cusim)"""
import os, random, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_LIB = os.path.join(ROOT, "tests", "cusim", "_build", "libstarneig_sim.so")
CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import starneig_b200 as sn
from starneig_b200 import api, _lib
from oracle.oracle import Oracle
api._handle = _lib.load(%r)
ora = Oracle()
sys.path.insert(0, %r)
from conftest import same_zero_pattern, structured_input
n, pw, gpus, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
entrywise = True
if kind == "fullpos":
    A0, Q0, ld = ora.fullpos(n, 2019)
else:       # tau = 0 columns, AED deflation window (tests/conftest.py)
    A0, Q0, ld, entrywise = structured_input(ora, kind, n)
A, Q = A0.copy(order="F"), Q0.copy(order="F")
sn.starneig_node_init(-1, gpus, sn.STARNEIG_NO_MESSAGES)
conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = pw
assert sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld) == 0
sn.starneig_node_finalize()
A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw)
u = 2.0 ** -52
# entrywise: 200 n u on the reference driver's fullpos matrices. The structured inputs start from the driver's `full` generator
# (entries of both signs) and are less well conditioned: two CPU reductions of the same input (the port at panel widths 8 and 35,
# zero_columns, n = 335) already differ by 175 n u in Q, so the bound is wider there; the invariants below stay at 500 u.
tol = (200 if kind == "fullpos" else 1000) * n * u
if entrywise:
    assert np.abs(A[:n] - A2[:n]).max() <= tol * max(1.0, np.abs(A2[:n]).max())
    assert np.abs(Q[:n] - Q2[:n]).max() <= tol
assert same_zero_pattern(A, A2, n)
assert ora.hessenberg_form_violations(n, A, ld) == 0
assert (not np.any(A0[:n]) or ora.residual_u(n, Q, ld, A, ld, A0, ld) <= 500) and ora.orthogonality_u(n, Q, ld) <= 500
print("OK")
''' % (ROOT, SIM_LIB, os.path.join(ROOT, "tests"))

KINDS = ["zero", "identity", "upper_triangular", "already_hessenberg", "zero_columns", "block_triangular", "aed_spike"]
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 240.0)
bad = runs = 0
while time.time() < t_end:
    P = random.choice([1, 1, 1, 2, 2, 3, 4, 8])
    n = random.choice([random.randint(3, 170), random.randint(260, 340)])
    pw = random.choice([8, 16, 24, 35, 64, 100])
    env = dict(os.environ, CUSIM_SMS=str(random.choice([1, 2, 3, 4, 6])), CUSIM_DEVICES="8", CUSIM_CHECK_PREFETCH="1",
               STARNEIG_B200_COL_BLOCK=str(random.choice([8, 16, 24])))
    sw = {}
    if random.random() < 0.3: sw["GEMV_LINEAR"] = 0
    if random.random() < 0.3: sw["GEMM_TMA"] = random.choice([0, 1, 2])
    if random.random() < 0.4: sw["GEMV_KC"] = random.choice([64, 128, 2048])
    if random.random() < 0.15: sw["FUSED_PANEL"] = 0
    if random.random() < 0.25: sw["Q_BACKWARD"] = 0          # the reference's forward order (default: backward for Q = I)
    if random.random() < 0.2: sw["FUSED_SLABS"] = 0
    if random.random() < 0.3: env["CUSIM_SHUFFLE"] = str(random.randint(1, 99))
    if random.random() < 0.3: env["CUSIM_SKEW"] = str(random.choice([2, 3, 5]))
    for k, v in sw.items():
        env["STARNEIG_B200_" + k] = str(v)
    kind = random.choice(["fullpos", "fullpos"] + KINDS) if n >= 12 else "fullpos"
    r = subprocess.run([sys.executable, "-c", CHILD, str(n), str(pw), str(P), kind], env=env, capture_output=True, text=True, timeout=900)
    runs += 1
    if r.returncode != 0 or not r.stdout.strip().endswith("OK"):
        bad += 1
        print("FAIL", dict(P=P, n=n, pw=pw, kind=kind), sw, {k: v for k, v in env.items() if k.startswith("CUSIM") or k == "STARNEIG_B200_COL_BLOCK"}, r.stderr[-400:], flush=True)
print("runs", runs, "failures", bad)
sys.exit(1 if bad else 0)
