"""CPU tests of the oracle itself (no GPU): the restatement ("port") against
 (i) golden fixtures produced by the reference's own sources (tests/golden/make_golden.py),
 (ii) the live reference build oracle/_ref when present, (iii) LAPACK dgehrd/dormhr (the reference test
 driver's `lapack` solver) and (iv) the reference driver's invariants (test/common/hooks.c:52-57,434-456).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_cases

U = 2.0 ** -52


def _inputs(ora, gen, n, begin, end, seed):
    if gen == "fullpos":
        return ora.fullpos(n, seed)
    if gen == "full":
        return ora.full(n, seed)
    return ora.partial(n, begin, end, seed)


def test_prand_known_answers(ora):
    # LCG of test/common/common.c:48-59; fixture generated together with the reference outputs
    seq = np.load(os.path.join(GOLDEN_DIR, "prand_seed2019.npz"))["seq"]
    ora.prand_init(2019)
    got = [ora.prand() for _ in range(len(seq))]
    assert got == list(seq)
    # closed form of the first step
    assert seq[0] == (2019 * 1103515245 + 12345) & 0x7FFFFFFF


def test_default_panel_width(ora, sn):
    # n=2000 -> 280, 10000 -> 296, 20000 -> 312, 50000 -> 368 (interface.c:74-78)
    for n, w in [(2000, 280), (10000, 296), (20000, 312), (50000, 368), (10, 280)]:
        assert ora.default_panel_width(n) == w
        assert sn.default_panel_width(n) == w


@pytest.mark.parametrize("case", golden_cases())
def test_port_matches_reference_golden(ora, case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    n, begin, end, pw = int(g["n"]), int(g["begin"]), int(g["end"]), int(g["panel_width"])
    A, Q, ld = _inputs(ora, str(g["generator"]), n, begin, end, int(g["seed"]))
    assert np.array_equal(A[:n], g["A0"]), "input generator drifted from the fixture"
    assert ora.hessenberg_port(n, A, ld, Q, ld, begin, end, pw) == 0
    scale = np.abs(g["H"]).max()
    assert np.abs(A[:n] - g["H"]).max() <= 200 * n * U * scale
    assert np.abs(Q[:n] - g["Q"]).max() <= 200 * n * U
    # exact zeros exactly where the reference has them
    assert np.array_equal(A[:n] == 0.0, g["H"] == 0.0)


@pytest.mark.parametrize("n,tile,pw", [(60, 16, 8), (200, 48, 45), (333, 197, 35), (554, 340, 170), (400, -1, -1)])
def test_port_matches_live_reference(ora, ref, n, tile, pw):
    A0, Q0, ld = ora.fullpos(n, 5)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ref.hessenberg_expert(n, A1, ld, Q1, ld, 0, n, tile, pw) == 0
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    scale = np.abs(A1[:n]).max()
    assert np.abs(A1[:n] - A2[:n]).max() <= 200 * n * U * scale
    assert np.abs(Q1[:n] - Q2[:n]).max() <= 200 * n * U
    assert ora.hessenberg_form_violations(n, A1, ld) == 0


@pytest.mark.parametrize("n,tile,pw,executors,begin,end", [
    (60, 16, 8, 3, 0, None), (200, 48, 45, 8, 0, None), (333, 64, 35, 2, 0, None), (554, 96, 170, 5, 0, None),
    (400, 56, -1, 1, 0, None), (333, 24, 16, 4, 83, 249), (1000, -1, -1, 8, 0, None)])
def test_reference_task_graph_in_parallel(ora, ref, n, tile, pw, executors, begin, end):
    """The StarPU stand-in's parallel schedule (oracle/ref_shim/mini_starpu.c: worker threads, dependencies inferred from the
    access modes in insertion order) against its inline schedule on the reference's own task graph: with sequential BLAS the
    order of every floating-point sum is the same, so H and Q agree bit for bit. bench.py --impl reference times the parallel
    schedule; this is what makes that number a number of the reference's algorithm."""
    end = n if end is None else end
    A0, Q0, ld = ora.partial(n, begin, end, 9) if (begin, end) != (0, n) else ora.fullpos(n, 5)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ref.set_workers(executors)
    try:
        assert ref.hessenberg_expert(n, A1, ld, Q1, ld, begin, end, tile, pw) == 0
        ref.tasks_executed(reset=True)
        ref.set_executors(executors)
        assert ref.hessenberg_expert(n, A2, ld, Q2, ld, begin, end, tile, pw) == 0
        tasks = ref.tasks_executed(reset=True)
    finally:
        ref.set_executors(0)
        ref.set_workers(1)
    assert tasks > (end - begin - 1) * 3               # at least prepare / compute / finish per column
    assert np.array_equal(A1, A2) and np.array_equal(Q1, Q2)
    assert ora.hessenberg_form_violations(n, A2, ld, begin, end, check_outside=True) == 0
    assert ora.residual_u(n, Q2, ld, A2, ld, A0, ld) < 500 and ora.orthogonality_u(n, Q2, ld) < 500


@pytest.mark.parametrize("n", [47, 88, 333])
def test_live_reference_partial(ora, ref, n):
    # partial-hessenberg ctest sizes (test/CMakeLists.txt:389-406): begin = n/4, end = 3n/4
    begin, end = n // 4, 3 * n // 4
    A0, Q0, ld = ora.partial(n, begin, end, 9)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ref.hessenberg_expert(n, A1, ld, Q1, ld, begin, end, 24, 16) == 0
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end, 16) == 0
    assert ora.hessenberg_form_violations(n, A1, ld, begin, end, check_outside=True) == 0
    assert ora.hessenberg_form_violations(n, A2, ld, begin, end, check_outside=True) == 0
    assert np.abs(A1[:n] - A2[:n]).max() <= 200 * n * U * np.abs(A1[:n]).max()
    assert np.abs(Q1[:n] - Q2[:n]).max() <= 200 * n * U
    assert ora.residual_u(n, Q1, ld, A1, ld, A0, ld) < 1000       # test/misc/partial_hessenberg.c:49


@pytest.mark.parametrize("n,pw", [(1, 8), (2, 8), (3, 8), (47, 8), (88, 35), (333, 45), (554, 170), (700, -1)])
def test_port_invariants_and_lapack(ora, n, pw):
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A, ld, Q, ld, 0, n, pw) == 0
    assert ora.hessenberg_form_violations(n, A, ld) == 0                      # hooks.c:442-444
    assert ora.residual_u(n, Q, ld, A, ld, A0, ld) < 500                      # warn threshold hooks.c:52
    assert ora.orthogonality_u(n, Q, ld) < 500
    # same DLARFG sign convention as LAPACK => entrywise agreement, no sign fix-up
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_lapack(n, A2, ld, Q2, ld) == 0
    assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * max(1.0, np.abs(A2[:n]).max())
    assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U


def test_port_invalid_panel_width(ora):
    A, Q, ld = ora.fullpos(20)
    assert ora.hessenberg_port(20, A, ld, Q, ld, 0, 20, 4) == 3       # STARNEIG_INVALID_CONFIGURATION


def test_port_nonidentity_q(ora):
    n = 120
    A0, _, ld = ora.full(n, 4)
    rng = np.random.default_rng(0)
    Qr, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Q0 = np.zeros((ld, n), order="F"); Q0[:n] = Qr
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A, ld, Q, ld, 0, n, 16) == 0
    # A0 = U H U^T and Q = Q0 U  =>  Q0^T Q H (Q0^T Q)^T = A0
    Uq = np.asfortranarray(np.vstack([Qr.T @ Q[:n], np.zeros((ld - n, n))]))
    assert ora.residual_u(n, Uq, ld, A, ld, A0, ld) < 500
    assert ora.orthogonality_u(n, Q, ld) < 500


@pytest.mark.parametrize("e", [600, -600])
def test_port_is_scale_invariant(ora, e):
    # dlarfg_ (dnrm2 + dlapy2) never squares an entry unscaled: scaling A by a power of two scales H exactly
    n, pw = 200, 35
    A0, Q0, ld = ora.full(n, 7)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A1, ld, Q1, ld, 0, n, pw) == 0
    s = 2.0 ** e
    A2, Q2 = (A0 * s).copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    assert np.isfinite(A2).all()
    assert np.array_equal(A2 / s, A1) and np.array_equal(Q2, Q1)


def test_eigenvalues_preserved(ora):
    n = 150
    A0, Q0, ld = ora.full(n, 8)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    ora.hessenberg_port(n, A, ld, Q, ld)
    ev_h = np.sort_complex(ora.eigenvalues(n, A, ld))
    ev_a = np.sort_complex(np.linalg.eigvals(A0[:n]))
    assert np.abs(ev_h - ev_a).max() <= 1e-10 * np.linalg.norm(A0[:n])
