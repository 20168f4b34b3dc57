"""Randomised runs of the kernel-logic emulator build (tests/cusim) under AddressSanitizer or UndefinedBehaviorSanitizer: the
product's host code AND its kernels (as emulated CUDA threads) with a red zone right behind every device allocation
(CUSIM_EXACT_ALLOC=1), random sizes / panel widths / rank counts / engine switches, each reduction in its own process and
checked against the oracle's invariants.   usage: cusim_sanitize.py asan|ubsan [seed] [seconds]   (builds the library first)"""
import os, random, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kind = sys.argv[1] if len(sys.argv) > 1 else "asan"
assert kind in ("asan", "ubsan")
subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cusim"), kind], check=True, capture_output=True)
LIB = os.path.join(ROOT, "tests", "cusim", "_build", f"libstarneig_sim_{kind}.so")
RUNTIME = subprocess.run(["/usr/bin/gcc", f"-print-file-name=lib{kind}.so"], capture_output=True, text=True).stdout.strip()
CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import starneig_b200 as sn
from starneig_b200 import api, _lib
from oracle.oracle import Oracle
api._handle = _lib.load(%r)
ora = Oracle()
n, pw, gpus = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A0, Q0, ld = ora.fullpos(n, 2019)
A, Q = A0.copy(order="F"), Q0.copy(order="F")
sn.starneig_node_init(-1, gpus, sn.STARNEIG_NO_MESSAGES)
conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = pw
assert sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld) == 0
sn.starneig_node_finalize()
assert ora.hessenberg_form_violations(n, A, ld) == 0
assert ora.residual_u(n, Q, ld, A, ld, A0, ld) <= 500 and ora.orthogonality_u(n, Q, ld) <= 500
print("OK")
''' % (ROOT, LIB)

random.seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t_end = time.time() + (float(sys.argv[3]) if len(sys.argv) > 3 else 300.0)
runs = bad = 0
while time.time() < t_end:
    P = random.choice([1, 1, 2, 3, 4, 8])
    n = random.choice([random.randint(3, 170), random.randint(260, 330)])
    pw = random.choice([8, 16, 24, 35, 64, 100])
    env = dict(os.environ, LD_PRELOAD=RUNTIME, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0",
               UBSAN_OPTIONS="print_stacktrace=1", CUSIM_EXACT_ALLOC="1", CUSIM_SMS=str(random.choice([1, 2, 3, 4, 6])),
               CUSIM_DEVICES="8", STARNEIG_B200_COL_BLOCK=str(random.choice([8, 16, 24])))
    sw = {}
    if random.random() < 0.6: sw["Q_BACKWARD"] = 1          # (threshold 1: backward accumulation at these small sizes too)
    if random.random() < 0.3: sw["GEMV_LINEAR"] = 0
    if random.random() < 0.3: sw["GEMM_TMA"] = random.choice([0, 1, 2])
    if random.random() < 0.4: sw["GEMV_KC"] = random.choice([64, 128, 2048])
    if random.random() < 0.15: sw["FUSED_PANEL"] = 0
    if random.random() < 0.2: sw["FUSED_SLABS"] = 0
    for k, v in sw.items():
        env["STARNEIG_B200_" + k] = str(v)
    r = subprocess.run([sys.executable, "-c", CHILD, str(n), str(pw), str(P)], env=env, capture_output=True, text=True, timeout=1800)
    runs += 1
    if r.returncode != 0 or not r.stdout.strip().endswith("OK"):
        bad += 1
        print("FAIL", dict(P=P, n=n, pw=pw), sw, r.stderr[-1500:], flush=True)
print(kind, "runs", runs, "failures", bad)
sys.exit(1 if bad else 0)
