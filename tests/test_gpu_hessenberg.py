"""GPU parity tests of the whole path through the reference-facing C ABI
(starneig_node_init / starneig_SEP_SM_Hessenberg[_expert], host buffers in, host buffers out), against the
CPU oracle on the same reference-test-driver inputs, against the golden fixtures produced by the
reference's own sources, and -- at sizes the oracle cannot reach quickly -- through the invariants the
reference test driver checks (test/common/hooks.c:52-57,434-456; test/common/checks.c:180-208).

Tolerances (FP64, u = 2^-52):
  * entrywise H, Q vs oracle/reference: 200*n*u*max|H| resp. 200*n*u (summation order differs; the Householder
    sign convention is DLARFG's in both, so no sign fix-up is applied);
  * residual ||Q H Q^T - A||_F/||A||_F and orthogonality ||Q Q^T - I||_F/sqrt(n): <= 10*n*u (BASELINE.json) and
    <= 500 u (reference warn threshold);
  * eigenvalues: 1e-10 * ||A||_F (BASELINE.json).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, STRUCTURED, aed_window_check, golden_cases, same_zero_pattern, structured_input

pytestmark = pytest.mark.gpu
U = 2.0 ** -52


def _run(sn, n, A, ld, Q, begin=0, end=None, pw=-1, tile=-1):
    conf = sn.starneig_hessenberg_init_conf()
    conf.panel_width = pw
    conf.tile_size = tile
    return sn.starneig_SEP_SM_Hessenberg_expert(conf, n, begin, n if end is None else end, A, ld, Q, ld)


def _check_invariants(ora, n, A, Q, A0, ld, begin=0, end=None, outside=False):
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert ora.hessenberg_form_violations(n, A, ld, begin, end, check_outside=outside) == 0
    res = ora.residual_u(n, Q, ld, A, ld, A0, ld)
    orth = ora.orthogonality_u(n, Q, ld)
    assert res <= max(10.0 * n, 20.0) and res <= 500, res
    assert orth <= max(10.0 * n, 20.0) and orth <= 500, orth


def _check_entrywise(n, A, Q, Aref, Qref):
    assert np.abs(A[:n] - Aref[:n]).max() <= 200 * n * U * max(1.0, np.abs(Aref[:n]).max())
    assert np.abs(Q[:n] - Qref[:n]).max() <= 200 * n * U
    assert same_zero_pattern(A, Aref, n)                          # same exact zeros (tests/conftest.py)


@pytest.mark.parametrize("case", golden_cases())
def test_against_reference_golden(node, ora, case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    n, begin, end = int(g["n"]), int(g["begin"]), int(g["end"])
    gen, seed = str(g["generator"]), int(g["seed"])
    A, Q, ld = (ora.fullpos(n, seed) if gen == "fullpos" else ora.full(n, seed) if gen == "full"
                else ora.partial(n, begin, end, seed))
    assert np.array_equal(A[:n], g["A0"])
    assert _run(node, n, A, ld, Q, begin, end, int(g["panel_width"]), int(g["tile_size"])) == 0
    _check_entrywise(n, A, Q, g["H"], g["Q"])


# sizes of the reference ctest matrix (test/CMakeLists.txt:366-406) and the degenerate ones
@pytest.mark.parametrize("n,pw", [(1, 8), (2, 8), (3, 8), (9, 8), (17, 8), (47, 16), (88, 35), (100, 100),
                                  (333, 45), (554, 170), (1000, 314), (1500, 400), (2000, -1)])
def test_against_oracle(node, ora, n, pw):
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, pw=pw) == 0
    _check_invariants(ora, n, A, Q, A0, ld)
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    _check_entrywise(n, A, Q, A2, Q2)


@pytest.mark.parametrize("n", [47, 88, 333, 554])
def test_partial_reduction(node, ora, n):
    begin, end = n // 4, 3 * n // 4
    A0, Q0, ld = ora.partial(n, begin, end, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, begin, end, pw=16) == 0
    _check_invariants(ora, n, A, Q, A0, ld, begin, end, outside=True)
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end, 16)
    _check_entrywise(n, A, Q, A2, Q2)


# x = 0 in DLARFG (tau = 0, H = I: reference src/hessenberg/cpu.c:140) in every / some columns, and the deflation
# window the Schur stage's AED step hands to the Hessenberg reduction (src/schur/core.c:893-929); see tests/conftest.py
@pytest.mark.parametrize("name", STRUCTURED)
@pytest.mark.parametrize("n,pw,end", [(300, 45, 300), (333, 64, 250)])
def test_structured_inputs(node, ora, name, n, pw, end):
    A0, Q0, ld, entrywise = structured_input(ora, name, n)
    A0[end:n, :end] = 0.0       # a partial reduction is a similarity only if nothing lies below the reduced block
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, 0, end, pw=pw) == 0
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert np.count_nonzero(np.tril(A[:end, :end], -2)) == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, end, pw) == 0
    if entrywise:
        _check_entrywise(n, A, Q, A2, Q2)
    else:
        assert same_zero_pattern(A, A2, n)
    if np.any(A0[:n]):
        res = ora.residual_u(n, Q, ld, A, ld, A0, ld)
        assert res <= 500, res
    orth = ora.orthogonality_u(n, Q, ld)
    assert orth <= 500, orth
    if name in ("zero", "identity", "upper_triangular", "already_hessenberg"):
        assert np.array_equal(A, A0)            # nothing to do: the matrix comes back bit for bit


def test_aed_window_with_general_q(node, ora):
    aed_window_check(node, ora, 400, 300, 64)


def test_simple_interface_wide_ld_and_general_q(node, ora):
    # examples/sep_sm_full_chain.c:63-75: A in [-1,1], ld = (n/8+1)*8; Q a general orthogonal matrix
    n = 301
    ld = (n // 8 + 1) * 8 + 24
    A0, _, _ = ora.full(n, 5, ld=ld)
    Qr, _ = np.linalg.qr(np.random.default_rng(1).standard_normal((n, n)))
    Q0 = np.zeros((ld, n), order="F"); Q0[:n] = Qr
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    assert np.array_equal(A[n:], A0[n:]) and np.array_equal(Q[n:], Q0[n:])      # padding rows untouched
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.hessenberg_port(n, A2, ld, Q2, ld)
    _check_entrywise(n, A, Q, A2, Q2)
    assert ora.orthogonality_u(n, Q, ld) < 500


def test_bitwise_reproducible(node, ora):
    n = 700
    A0, Q0, ld = ora.fullpos(n, 3)
    outs = []
    for _ in range(2):
        A, Q = A0.copy(order="F"), Q0.copy(order="F")
        assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
        outs.append((A, Q))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_pinned_and_pageable_hosts_agree(node, ora):
    import torch
    n = 400
    A0, Q0, ld = ora.fullpos(n, 4)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    tp = torch.empty((2, n, ld), dtype=torch.float64).pin_memory()
    Ap = tp[0].numpy().T; Qp = tp[1].numpy().T                 # column-major (ld x n) views of pinned memory
    Ap[...] = A0; Qp[...] = Q0
    node.starneig_node_disable_pinning()
    try:
        assert node.starneig_SEP_SM_Hessenberg(n, Ap, ld, Qp, ld) == 0
    finally:
        node.starneig_node_enable_pinning()
    assert np.array_equal(Ap, A) and np.array_equal(Qp, Q)


def test_device_resident_matches_host_api(node, ora):
    import torch
    n = 600
    A0, Q0, ld = ora.fullpos(n, 6)
    ld = (n + 15) // 16 * 16
    A0, Q0, ld = ora.fullpos(n, 6, ld=ld)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    Ad = torch.from_numpy(np.ascontiguousarray(A0.T)).cuda()
    Qd = torch.from_numpy(np.ascontiguousarray(Q0.T)).cuda()
    assert node.hessenberg_device(n, Ad, ld, Qd, ld) == 0
    assert np.array_equal(Ad.cpu().numpy().T[:n], A[:n]) and np.array_equal(Qd.cpu().numpy().T[:n], Q[:n])
    st = node.get_stats()
    assert st["kernel_launches"] > 0 and st["gemv_launches"] == n - 1
    # unaligned / odd-ld device buffers are rejected, not silently mishandled
    assert node.lib().starneig_b200_hessenberg_device(n, 0, n, -1, Ad.data_ptr() + 8, ld, Qd.data_ptr(), ld) == -6
    assert node.lib().starneig_b200_hessenberg_device(n, 0, n, -1, Ad.data_ptr(), ld, Qd.data_ptr(), n + 1) in (-8,)


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("e", [600, -600])
def test_extreme_scaling(node, ora, monkeypatch, e, fused):
    # The reference forms its reflectors with LAPACK dlarfg_ (src/hessenberg/cpu.c:140), whose dnrm2 / dlapy2 neither
    # overflow nor underflow: a matrix scaled by 2^+-600 (squares of its entries are inf / 0 in FP64) reduces to the
    # scaled H and the same Q. Both panel paths: the persistent kernel and the three-kernels-per-column one.
    monkeypatch.setenv("STARNEIG_B200_FUSED_PANEL", str(fused))
    n, pw = 333, 45
    A0, Q0, ld = ora.full(n, 7)
    s = 2.0 ** e
    A, Q = (A0 * s).copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, pw=pw) == 0
    A2, Q2 = (A0 * s).copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    A /= s
    A2 /= s          # exact: s is a power of two
    _check_entrywise(n, A, Q, A2, Q2)
    _check_invariants(ora, n, A, Q, A0, ld)


def test_downstream_eigenvalues(node, ora):
    # config 5 of BASELINE.json at test size: GPU Hessenberg -> dhseqr (stand-in for starneig_SEP_SM_Schur)
    # vs the all-CPU chain (oracle Hessenberg -> dhseqr); tolerance 1e-10 * ||A||
    n = 800
    A0, Q0, ld = ora.full(n, 12)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.hessenberg_port(n, A2, ld, Q2, ld)
    ev_gpu = np.sort_complex(ora.eigenvalues(n, A, ld))
    ev_cpu = np.sort_complex(ora.eigenvalues(n, A2, ld))
    norm = np.linalg.norm(A0[:n])
    # match each GPU eigenvalue with the nearest CPU eigenvalue (sorting can swap near-equal real parts)
    d = np.abs(ev_gpu[:, None] - ev_cpu[None, :]).min(axis=1)
    assert d.max() <= 1e-10 * norm


@pytest.mark.parametrize("n", [4000])
def test_invariants_at_larger_size(node, ora, n):
    # size-independent properties: exact-zero Hessenberg form, residual, orthogonality, trace preservation
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    _check_invariants(ora, n, A, Q, A0, ld)
    assert abs(np.trace(A[:n]) - np.trace(A0[:n])) <= 100 * n * U * abs(np.trace(A0[:n]))
    assert abs(np.linalg.norm(A[:n]) - np.linalg.norm(A0[:n])) <= 100 * n * U * np.linalg.norm(A0[:n])


# the largest case of the reference's ctest matrix (test/CMakeLists.txt:366-406: n = 3569, odd, panel widths 303 / 410)
@pytest.mark.parametrize("n,pw", [(3569, 410)])
def test_invariants_reference_ctest_size(node, ora, n, pw):
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, pw=pw) == 0
    _check_invariants(ora, n, A, Q, A0, ld)


# ---------------------------------------------------------------------------------------------------------------------
# The sizes that carry the benchmark numbers. The acceptance checks are the reference driver's (tools/invariants.py),
# evaluated on the GPU with torch FP64 matmuls so that n = 10000 takes seconds; the reduction itself goes through the
# reference-facing C call on host buffers.
# ---------------------------------------------------------------------------------------------------------------------
def _gpu_invariants(n, A, Q, A0, ld, begin=0, end=None):
    import torch
    from tools import invariants
    t = lambda M: torch.from_numpy(np.ascontiguousarray(M.T)).cuda()       # (n, ld): row c = column c
    out = invariants.evaluate(t(A0), t(A), t(Q), n, begin, end)
    torch.cuda.empty_cache()
    return out


# the reference's ctest grid for this path: n = 4000 x every panel width (test/CMakeLists.txt:367,384-389)
@pytest.mark.parametrize("pw", [45, 314, 400, 410, 170, 35, 303])
def test_reference_ctest_panel_widths_n4000(node, ora, pw):
    n = 4000
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, pw=pw) == 0
    inv = _gpu_invariants(n, A, Q, A0, ld)
    assert inv["ok"], inv


# ... and its tile sizes (test/CMakeLists.txt:366,377-382): accepted, validated and ignored by this engine (no tiles), so
# every one of them must give the bits of the default call
@pytest.mark.parametrize("ts", [48, 549, 611, 883, 448, 340, 526, 197])
def test_reference_ctest_tile_sizes_are_accepted(node, ora, ts):
    n = 1200
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, tile=ts) == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A2, ld, Q2, ld) == 0
    assert np.array_equal(A, A2) and np.array_equal(Q, Q2)


def test_entrywise_against_oracle_n4000(node, ora):
    n = 4000
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.set_threads(os.cpu_count() or 1)
    assert ora.hessenberg_port(n, A2, ld, Q2, ld) == 0
    _check_entrywise(n, A, Q, A2, Q2)


# the largest partial reduction of the reference's ctest matrix (test/CMakeLists.txt:391-399: n = 3569, begin = n/4,
# end = 3n/4)
def test_partial_reduction_reference_ctest_size(node, ora):
    n = 3569
    begin, end = n // 4, 3 * n // 4
    A0, Q0, ld = ora.partial(n, begin, end, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(node, n, A, ld, Q, begin, end) == 0
    assert ora.hessenberg_form_violations(n, A, ld, begin, end, check_outside=True) == 0
    inv = _gpu_invariants(n, A, Q, A0, ld, begin, end)
    assert inv["ok"], inv
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    ora.set_threads(os.cpu_count() or 1)
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end) == 0
    _check_entrywise(n, A, Q, A2, Q2)


# BASELINE.json configs[1]: random dense n = 10000 with Q on one B200
def test_invariants_n10000(node, ora):
    n = 10000
    rng = np.random.default_rng(2019)
    ld = n
    A0 = np.asfortranarray(rng.random((n, n)))
    Q0 = np.asfortranarray(np.eye(n))
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    inv = _gpu_invariants(n, A, Q, A0, ld)
    assert inv["ok"] and inv["trace_rel_err"] <= 100 * n * U, inv
    print("n=10000 invariants:", inv)


# ---------------------------------------------------------------------------------------------------------------------
# chain hand-off
# ---------------------------------------------------------------------------------------------------------------------
def _device_to_host(ptr, ldd, n):
    """(ldd x n) column-major device matrix -> numpy, through the CUDA runtime (no product code involved)"""
    from cuda import cudart
    out = np.zeros((ldd, n), order="F")
    err, = cudart.cudaMemcpy(out.ctypes.data, ptr, ldd * n * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    return out


def test_hessenberg_stage_leaves_h_and_q_on_the_device(node, ora):
    n = 700
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert node.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    ret, dH, lddH, dQ, lddQ = node.hessenberg_stage(n, A1, ld, Q1, ld)
    assert ret == 0 and np.array_equal(A1, A0) and np.array_equal(Q1, Q0) and node.get_stats()["d2h_bytes"] == 0
    H = _device_to_host(dH, lddH, n)
    Qd = _device_to_host(dQ, lddQ, n)
    assert np.array_equal(H[:n], A[:n]) and np.array_equal(Qd[:n], Q[:n])       # the device copies ARE the result
    assert node.stage_fetch(n, A1, ld, Q1, ld) == 0
    assert np.array_equal(A1[:n], A[:n]) and np.array_equal(Q1[:n], Q[:n])


def test_reduce_shaped_chain_with_a_device_schur_stage(node, ora):
    """starneig_b200_SEP_SM_Reduce (shape of reference src/common/combined.c:45-98): the next stage receives DEVICE pointers;
    the stand-in for the Schur stage pulls H from the device itself and takes its eigenvalues with LAPACK dhseqr"""
    import ctypes
    from starneig_b200 import Chain, SCHUR_FN
    n = 600
    A0, Q0, ld = ora.full(n, 12)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    real, imag = np.zeros(n), np.zeros(n)
    seen = []

    def schur_device(nn, pH, ldH, pQ, ldQ, preal, pimag):
        H = _device_to_host(pH, ldH, nn)
        ev = ora.eigenvalues(nn, H, ldH)
        np.ctypeslib.as_array(ctypes.cast(preal, ctypes.POINTER(ctypes.c_double)), shape=(nn,))[:] = ev.real
        np.ctypeslib.as_array(ctypes.cast(pimag, ctypes.POINTER(ctypes.c_double)), shape=(nn,))[:] = ev.imag
        seen.append((ldH, ldQ))
        return 0

    chain = Chain()
    fn = SCHUR_FN(schur_device)
    chain.schur_device = fn
    ret, _ = node.starneig_b200_SEP_SM_Reduce(n, A, ld, Q, ld, real, imag, chain)
    assert ret == 0 and len(seen) == 1
    ev = real + 1j * imag
    want = np.linalg.eigvals(A0[:n])
    d = np.abs(ev[:, None] - want[None, :]).min(axis=1)
    assert d.max() <= 1e-10 * np.linalg.norm(A0[:n])
    _check_invariants(ora, n, A, Q, A0, ld)         # H and Q came back through the fetch at the end of the chain
