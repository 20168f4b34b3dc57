"""GPU tests of the engine's switches and rare branches, each in its own process with a time-out (a fault or a hang fails
one test instead of taking the whole GPU suite with it): the DLARFG rescaling branch, GEMV linearity against the
sequential order of operations, and the switch that must not change a single bit (staging chunk of the GEMV). All of it is also stepped through the kernel-logic emulator
(tests/test_cusim.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import starneig_b200 as sn
from oracle.oracle import Oracle
ora = Oracle()
U = 2.0 ** -52
mode, n, pw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])

def run(A0, Q0, ld, env):
    for k, v in env.items():
        os.environ[k] = v
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    try:
        conf = sn.starneig_hessenberg_init_conf()
        conf.panel_width = pw
        for _ in range(2):          # twice: the LL tags of the second call continue where the first one stopped
            A[:], Q[:] = A0, Q0
            assert sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld) == 0
        run.stats = sn.get_stats()
    finally:
        sn.starneig_node_finalize()
    for k in env:
        os.environ.pop(k, None)
    return A, Q

if mode == "denormal":
    # A matrix scaled by 2^-1040 lives in the denormal range: LAPACK's dlarfg (reference src/hessenberg/cpu.c:140)
    # rescales x and alpha by 1/safmin before it forms the reflector, and so do the kernels (dlarfg_scalars, panel.cuh).
    # The data carries only ~34 bits there, so H and Q agree with the reference to that level only -- but Q stays
    # orthogonal to a few u and nothing overflows (without the branch: 1 / (alpha - beta) = inf => NaN).
    A0, Q0, ld = ora.full(n, 7)
    A0 = (A0 * 2.0 ** -1040).copy(order="F")
    A, Q = run(A0, Q0, ld, {"STARNEIG_B200_FUSED_PANEL": sys.argv[4]})
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert ora.hessenberg_form_violations(n, A, ld) == 0
    assert ora.orthogonality_u(n, Q, ld) <= 500
    assert np.abs(A[:n] - A2[:n]).max() <= 1e-5 * np.abs(A2[:n]).max()
    assert np.abs(Q[:n] - Q2[:n]).max() <= 1e-5
elif mode == "sequential_gemv":
    A0, Q0, ld = ora.fullpos(n, 2019)
    env = {"STARNEIG_B200_GEMV_LINEAR": "0"}
    A, Q = run(A0, Q0, ld, env)
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * max(1.0, np.abs(A2[:n]).max())
    assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U
    assert ora.hessenberg_form_violations(n, A, ld) == 0
    assert ora.residual_u(n, Q, ld, A, ld, A0, ld) <= 500 and ora.orthogonality_u(n, Q, ld) <= 500
    if True:
        # ... and the default (linear) order of operations gives the same reduction up to where `scale` is applied
        A1, Q1 = run(A0, Q0, ld, {})
        assert np.abs(A[:n] - A1[:n]).max() <= 200 * n * U * max(1.0, np.abs(A1[:n]).max())
        assert np.abs(Q[:n] - Q1[:n]).max() <= 200 * n * U and not np.array_equal(A, A1)
elif mode == "q_forward":
    # Q = I on entry: the default accumulates Q backward after the last panel (engine.cuh, Rank::reduce); the forward order
    # (the reference's, STARNEIG_B200_Q_BACKWARD=0) gives the same H bit for bit and the same Q up to rounding
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = run(A0, Q0, ld, {})
    assert run.stats["q_backward"] == 1, run.stats
    A1, Q1 = run(A0, Q0, ld, {"STARNEIG_B200_Q_BACKWARD": "0"})
    assert run.stats["q_backward"] == 0
    assert np.array_equal(A, A1)
    assert np.abs(Q[:n] - Q1[:n]).max() <= 50 * n * U and not np.array_equal(Q, Q1)
    for (AA, QQ) in ((A, Q), (A1, Q1)):
        assert ora.hessenberg_form_violations(n, AA, ld) == 0
        assert ora.residual_u(n, QQ, ld, AA, ld, A0, ld) <= 500 and ora.orthogonality_u(n, QQ, ld) <= 500
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U
else:
    # opt-in variant against the default kernels: same partial sums in the same order => bitwise the same H and Q
    key, val = sys.argv[4].split("=")
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = run(A0, Q0, ld, {})
    A1, Q1 = run(A0, Q0, ld, {key: val})
    assert np.array_equal(A1[:n], A[:n]) and np.array_equal(Q1[:n], Q[:n])
    assert ora.hessenberg_form_violations(n, A1, ld) == 0
    assert ora.residual_u(n, Q1, ld, A1, ld, A0, ld) <= 500 and ora.orthogonality_u(n, Q1, ld) <= 500
print("OK")
''' % ROOT


def _child(*args, timeout=180):
    r = subprocess.run([sys.executable, "-c", CHILD] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-1500:], r.stderr[-3000:])


@pytest.mark.parametrize("fused", [1, 0])
def test_denormal_range_takes_dlarfg_rescaling_branch(fused):
    _child("denormal", 333, 45, fused)


@pytest.mark.parametrize("switch", ["STARNEIG_B200_GEMV_KC=2048", "STARNEIG_B200_FUSED_SLABS=0"])
def test_switch_is_bitwise_equal_to_the_default(switch):
    _child("variant", 1500, 200, switch)


@pytest.mark.parametrize("n,pw", [(1500, 200), (2001, 192)])
def test_backward_accumulation_of_q_agrees_with_the_forward_order(n, pw):
    _child("q_forward", n, pw, 0)


def test_sequential_gemv_order_agrees_with_the_linear_default():
    _child("sequential_gemv", 1500, 200, 0)
