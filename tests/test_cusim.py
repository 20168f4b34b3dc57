"""Kernel-logic tests on a machine WITHOUT a GPU: the product's CUDA sources under the emulator of tests/cusim.

`tests/cusim` compiles the PRODUCT sources (starneig_b200/csrc/hessenberg.cu, engine.cuh, panel*.cuh, dgemm.cuh,
node.cpp) unchanged with g++ against a stand-in CUDA runtime: every CUDA thread is a fiber, block/named barriers, warp
shuffles, the m8n8k4 FP64 MMA fragment layout, cooperative grids, acquire/release flags and the 16-byte LL entries of the
multi-GPU exchange behave as the kernels assume, ranks are host threads that really run concurrently. What these tests
pin is therefore the LOGIC of the kernels and of the launch sequence (indexing, barrier structure, exchange protocol,
edge cases) -- through the same C ABI and against the same oracle as the GPU parity tests -- not their speed, and not the
GPU's memory model. It is test infrastructure: the product library never links or loads it
(tests/test_abi.py::test_product_never_touches_the_oracle), and the GPU tests (-m gpu) remain the parity gate.

Tolerances as in tests/test_gpu_hessenberg.py: entrywise 200*n*u, residual / orthogonality <= 10*n*u and <= 500 u.
"""
import os
import platform
import subprocess

import numpy as np
import pytest

from conftest import STRUCTURED, aed_window_check, same_zero_pattern, structured_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libstarneig_sim.so")
U = 2.0 ** -52

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's context switch is x86-64 only")


@pytest.fixture(scope="module")
def simlib():
    r = subprocess.run(["make", "-C", SIM_DIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    from starneig_b200 import _lib
    return _lib.load(SIM_LIB)


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture()
def sim(simlib, monkeypatch):
    """the host-side mirror (starneig_b200.api) bound to the emulator build for the duration of one test"""
    import starneig_b200
    from starneig_b200 import api
    monkeypatch.setattr(api, "_handle", simlib)
    yield starneig_b200
    if simlib.starneig_node_initialized():
        simlib.starneig_node_finalize()


def _reduce(sn, ora, n, pw, gpus=1, begin=0, end=None, generator="fullpos", ld_extra=0, given=None, entrywise=True):
    end = n if end is None else end
    if given is not None:
        A0, Q0, ld = given
    elif generator == "partial":
        A0, Q0, ld = ora.partial(n, begin, end, 2019)
    else:
        A0, Q0, ld = ora.fullpos(n, 2019)
    if ld_extra:
        ld2 = ld + ld_extra
        A0 = np.asfortranarray(np.vstack([A0, np.full((ld_extra, A0.shape[1]), np.nan)]))
        Q0 = np.asfortranarray(np.vstack([Q0, np.full((ld_extra, Q0.shape[1]), np.nan)]))
        ld = ld2
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, gpus, sn.STARNEIG_NO_MESSAGES)
    try:
        conf = sn.starneig_hessenberg_init_conf()
        conf.panel_width = pw
        ret = sn.starneig_SEP_SM_Hessenberg_expert(conf, n, begin, end, A, ld, Q, ld)
        stats = sn.get_stats()
    finally:
        sn.starneig_node_finalize()
    assert ret == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end, pw) == 0
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    if entrywise:
        assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * max(1.0, np.abs(A2[:n]).max())
        assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U
    assert same_zero_pattern(A, A2, n)
    if given is None:
        assert ora.hessenberg_form_violations(n, A, ld, begin, end, check_outside=(generator == "partial")) == 0
    res, orth = ora.residual_u(n, Q, ld, A, ld, A0, ld) if np.any(A0[:n]) else 0.0, ora.orthogonality_u(n, Q, ld)
    assert res <= max(10.0 * n, 20.0) and res <= 500 and orth <= max(10.0 * n, 20.0) and orth <= 500
    if ld_extra:
        assert np.isnan(A[ld - ld_extra:]).all() and np.isnan(Q[ld - ld_extra:]).all()      # padding rows stay untouched
    return A, Q, stats


# the persistent panel kernel + DMMA updates, one rank; degenerate sizes, odd sizes, panel wider than the matrix
@pytest.mark.parametrize("n,pw", [(1, 8), (2, 8), (3, 8), (9, 8), (17, 8), (47, 16), (88, 35), (100, 100), (131, 24)])
def test_sim_fused_single_rank(sim, ora, n, pw):
    _, _, st = _reduce(sim, ora, n, pw)
    assert st["fused_panels"] == st["panels"] and (n < 3 or st["kernel_launches"] > 0)


def test_sim_more_ctas_than_rows(sim, ora):
    with _Env(CUSIM_SMS=8):
        _reduce(sim, ora, 60, 16)


def test_sim_several_subtiles_per_cta(sim, ora):
    with _Env(CUSIM_SMS=1):
        _reduce(sim, ora, 90, 16)          # nsub = 3: one CTA owns three 32-row sub-tiles


# the three-kernels-per-column path (panels wider than FUSED_MAX_NB, devices without cooperative launch)
@pytest.mark.parametrize("n,pw", [(47, 16), (100, 40)])
def test_sim_unfused_single_rank(sim, ora, n, pw):
    with _Env(STARNEIG_B200_FUSED_PANEL=0):
        _, _, st = _reduce(sim, ora, n, pw)
    assert st["fused_panels"] == 0


@pytest.mark.parametrize("gpus,n,pw,sms", [(1, 131, 24, 4), (1, 90, 16, 1), (2, 96, 16, 2), (4, 150, 40, 2)])
def test_sim_linear_gemv_against_sequential(sim, ora, gpus, n, pw, sms):
    """GEMV linearity (the default; FusedArgs::linear): the GEMV streams against the unscaled x while the look-ahead warps
    derive the DLARFG scalars, y = A(:, c+1) + scale (A(:, c+2:) x) is formed by the row owners. Against the sequential
    order (STARNEIG_B200_GEMV_LINEAR=0: scalars first, then A v): the same reduction up to where `scale` is applied --
    both within the oracle tolerance (checked by _reduce), and within it of each other."""
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=sms):
        A, Q, st = _reduce(sim, ora, n, pw, gpus=gpus)
        with _Env(STARNEIG_B200_GEMV_LINEAR=0):
            A1, Q1, _ = _reduce(sim, ora, n, pw, gpus=gpus)
    assert st["fused_panels"] == st["panels"]
    assert np.abs(A[:n] - A1[:n]).max() <= 200 * n * U * np.abs(A1[:n]).max() and np.abs(Q[:n] - Q1[:n]).max() <= 200 * n * U
    assert same_zero_pattern(A, A1, n)
    assert not np.array_equal(A, A1)            # the switch does select another order of operations


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("e", [600, -600, -1040])
def test_sim_extreme_scaling(sim, ora, e, fused):
    """2^+-600: the squares of the entries overflow / underflow (Blue's sums of squares). 2^-1040: the matrix lives in
    the denormal range, LAPACK's dlarfg takes its rescaling branch (|beta| < safmin) and so must the kernels: Q stays
    orthogonal to a few u although the data itself carries only ~34 bits there."""
    n, pw = 60, 16
    A0, Q0, ld = ora.full(n, 7)
    s = 2.0 ** e
    A, Q = (A0 * s).copy(order="F"), Q0.copy(order="F")
    with _Env(STARNEIG_B200_FUSED_PANEL=fused):
        sim.starneig_node_init(sim.STARNEIG_USE_ALL, 1, sim.STARNEIG_NO_MESSAGES)
        try:
            conf = sim.starneig_hessenberg_init_conf()
            conf.panel_width = pw
            assert sim.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld) == 0
        finally:
            sim.starneig_node_finalize()
    A2, Q2 = (A0 * s).copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert ora.hessenberg_form_violations(n, A, ld) == 0
    assert ora.orthogonality_u(n, Q, ld) <= 50
    tol = 200 * n * U if e > -1000 else 1e-6          # denormal data: 2^-1040 leaves 34 significant bits
    assert np.abs(A[:n] - A2[:n]).max() <= tol * np.abs(A2[:n]).max()
    assert np.abs(Q[:n] - Q2[:n]).max() <= tol


@pytest.mark.parametrize("n", [47, 88])
def test_sim_partial_reduction(sim, ora, n):
    _reduce(sim, ora, n, 16, begin=n // 4, end=3 * n // 4, generator="partial")


# x = 0 in DLARFG (tau = 0) in every / some columns, and the AED window of the Schur stage (tests/conftest.py)
@pytest.mark.parametrize("name", STRUCTURED)
@pytest.mark.parametrize("gpus,n,pw,end,fused", [(1, 40, 16, 40, 1), (1, 70, 24, 52, 1), (2, 64, 16, 64, 1), (1, 40, 16, 30, 0)])
def test_sim_structured_inputs(sim, ora, name, gpus, n, pw, end, fused):
    A0, Q0, ld, entrywise = structured_input(ora, name, n)
    A0[end:n, :end] = 0.0       # a partial reduction is a similarity only if nothing lies below the reduced block
    with _Env(STARNEIG_B200_FUSED_PANEL=fused):
        A, _, _ = _reduce(sim, ora, n, pw, gpus=gpus, end=end, given=(A0, Q0, ld), entrywise=entrywise)
    assert np.count_nonzero(np.tril(A[:end, :end], -2)) == 0
    if name in ("zero", "identity", "upper_triangular", "already_hessenberg"):
        assert np.array_equal(A, A0)            # nothing to do: tau = 0 everywhere, the matrix comes back bit for bit


# the AED client accumulates into its local (non-identity) Q: Q <- Q0 U, so Q H Q^T = Q0 A0 Q0^T
@pytest.mark.parametrize("gpus,n,end,pw", [(1, 60, 45, 16), (2, 64, 48, 16)])
def test_sim_aed_window_with_general_q(sim, ora, gpus, n, end, pw):
    aed_window_check(sim, ora, n, end, pw, gpus)


# BASELINE.json configs[4] at emulator size: this path's H -> dhseqr (stand-in for starneig_SEP_SM_Schur) against the
# all-CPU chain (reference port -> dhseqr), eigenvalues within 1e-10 * ||A||
@pytest.mark.parametrize("gpus", [1, 2])
def test_sim_downstream_eigenvalues(sim, ora, gpus):
    n = 150
    A0, Q0, ld = ora.full(n, 12)
    A, _, _ = _reduce(sim, ora, n, 24, gpus=gpus, given=(A0, Q0, ld))
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, 24) == 0
    ev, ev_cpu = ora.eigenvalues(n, A, ld), ora.eigenvalues(n, A2, ld)
    d = np.abs(ev[:, None] - ev_cpu[None, :]).min(axis=1)
    assert d.max() <= 1e-10 * np.linalg.norm(A0[:n])


def test_sim_padded_leading_dimension(sim, ora):
    _reduce(sim, ora, 50, 16, ld_extra=6)


# multi-GPU engine: ranks are host threads with their own schedulers, exchanging through "peer" memory
@pytest.mark.parametrize("gpus,n,pw,cb", [(2, 96, 16, 8), (3, 70, 16, 8), (4, 120, 24, 16)])
def test_sim_multi_rank_fused(sim, ora, gpus, n, pw, cb):
    with _Env(STARNEIG_B200_COL_BLOCK=cb, CUSIM_SMS=2):
        _, _, st = _reduce(sim, ora, n, pw, gpus=gpus)
    assert st["ranks"] == gpus and st["fused_panels"] == st["panels"]


def test_sim_eight_ranks(sim, ora):
    """the full width of one NVSwitch box: 8 ranks, one (emulated) SM each, column blocks of 8"""
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=1, CUSIM_DEVICES=8):
        _, _, st = _reduce(sim, ora, 150, 24, gpus=8)
    assert st["ranks"] == 8 and st["fused_panels"] == st["panels"]


@pytest.mark.parametrize("gpus,n,cb,linear", [(4, 9, 8, 0), (4, 20, 16, 1), (2, 5, 8, 1), (8, 40, 8, 1)])
def test_sim_ranks_with_few_or_no_columns(sim, ora, gpus, n, cb, linear):
    """matrices smaller than one round of column blocks: some ranks own one block, some nothing at all"""
    with _Env(STARNEIG_B200_COL_BLOCK=cb, CUSIM_SMS=2, CUSIM_DEVICES=8, STARNEIG_B200_GEMV_LINEAR=linear):
        _, _, st = _reduce(sim, ora, n, 8, gpus=gpus)
    assert st["ranks"] == gpus


def test_sim_multi_rank_unfused(sim, ora):
    with _Env(STARNEIG_B200_COL_BLOCK=8, STARNEIG_B200_FUSED_PANEL=0):
        _, _, st = _reduce(sim, ora, 72, 16, gpus=2)
    assert st["ranks"] == 2 and st["fused_panels"] == 0


def test_sim_multi_rank_partial(sim, ora):
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=2):
        _reduce(sim, ora, 88, 16, gpus=2, begin=22, end=66, generator="partial")


def test_sim_results_do_not_depend_on_the_schedule(sim, ora, simlib):
    """fixed-order reductions: the result is bitwise the same whatever order the threads are resumed in"""
    A1, Q1, _ = _reduce(sim, ora, 64, 16)
    r = subprocess.run(["python", "-c", (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import starneig_b200 as sn\n"
        "from starneig_b200 import api, _lib\n"
        "from oracle.oracle import Oracle\n"
        "api._handle = _lib.load(%r)\n"
        "A, Q, ld = Oracle().fullpos(64, 2019)\n"
        "sn.starneig_node_init(-1, 1, sn.STARNEIG_NO_MESSAGES)\n"
        "conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = 16\n"
        "assert sn.starneig_SEP_SM_Hessenberg_expert(conf, 64, 0, 64, A, ld, Q, ld) == 0\n"
        "sn.starneig_node_finalize()\n"
        "np.save(sys.argv[1], np.stack([A[:64], Q[:64]]))\n") % (ROOT, SIM_LIB), "/tmp/cusim_shuffled.npy"],
        env=dict(os.environ, CUSIM_SHUFFLE="7"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    AQ = np.load("/tmp/cusim_shuffled.npy")
    assert np.array_equal(AQ[0], A1[:64]) and np.array_equal(AQ[1], Q1[:64])


@pytest.mark.parametrize("gpus,n,pw", [(1, 131, 24), (1, 90, 35), (2, 96, 16)])
def test_sim_all_updates_take_the_tma_kernels(sim, ora, gpus, n, pw):
    """Every DMMA launch of a reduction is framed for the TMA kernels (dgemm_tma.cuh): the emulator is as strict as the
    hardware about 16-byte box origins, panels start at odd and even rows, panel widths are odd and even -- only products
    without a k extent (a rank that owns no column of the range) fall back to the cp.async kernels."""
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=2):
        _, _, st = _reduce(sim, ora, n, pw, gpus=gpus)
    assert st["gemm_tma_launches"] > 0
    assert st["gemm_cpasync_launches"] == 0 or gpus > 1, st
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=2, STARNEIG_B200_GEMM_TMA=0):
        _, _, st = _reduce(sim, ora, n, pw, gpus=gpus)
    assert st["gemm_tma_launches"] == 0 and st["gemm_cpasync_launches"] > 0


def test_sim_dgemm_kinds(sim, simlib):
    """the three operand layouts of the DMMA kernel (fragment layout of mma.m8n8k4 emulated lane by lane), edges, odd
    sizes, split-K; operands at 16-byte aligned and at odd (8-byte aligned) offsets"""
    rng = np.random.default_rng(1)
    sim.starneig_node_init(sim.STARNEIG_USE_ALL, 1, sim.STARNEIG_NO_MESSAGES)
    try:
        for (ta, tb, m, n, k) in [("N", "T", 70, 37, 21), ("T", "N", 45, 13, 600), ("N", "N", 83, 29, 1100), ("N", "T", 130, 66, 4),
                                  ("N", "T", 64, 64, 48), ("T", "N", 129, 104, 40)]:
            for (offa, offb) in [(0, 0), (1, 0), (0, 1), (1, 1)]:
                # operands as sub-matrices starting at row offa / offb of 16-byte aligned, even-ld arrays
                ra, ca = ((m, k) if ta == "N" else (k, m))
                rb, cb = ((n, k) if tb == "T" else (k, n))
                lda, ldb, ldc = (ra + offa + 3) // 2 * 2, (rb + offb + 3) // 2 * 2, (m + 1) // 2 * 2
                abuf = np.asfortranarray(rng.standard_normal((lda, ca)))
                bbuf = np.asfortranarray(rng.standard_normal((ldb, cb)))
                c = np.asfortranarray(rng.standard_normal((ldc, n)))
                a, b = abuf[offa:offa + ra], bbuf[offb:offb + rb]
                beta = 1.0 if (ta, tb) == ("N", "T") else 0.0
                want = 0.75 * (a if ta == "N" else a.T) @ (b.T if tb == "T" else b) + beta * c[:m]
                ret = simlib.starneig_b200_dgemm(ta.encode(), tb.encode(), m, n, k, 0.75, abuf.ctypes.data + 8 * offa, lda,
                                                 bbuf.ctypes.data + 8 * offb, ldb, beta, c.ctypes.data, ldc)
                assert ret == 0
                assert np.abs(c[:m] - want).max() <= 50 * k * U * np.abs(want).max(), (ta, tb, m, n, k, offa, offb)
    finally:
        sim.starneig_node_finalize()


_PANEL_CHILD = r"""
import sys, os
sys.path.insert(0, %r)
import numpy as np
import starneig_b200 as sn
from starneig_b200 import api, _lib
from oracle.oracle import Oracle
lib = api._handle = _lib.load(%r)
n, w = int(sys.argv[1]), int(sys.argv[2])
A0, Q0, ld = Oracle().fullpos(n, 2019)
def panel(env):
    os.environ.update(env)
    sn.starneig_node_init(-1, 1, sn.STARNEIG_NO_MESSAGES)
    A = A0.copy(order="F")
    ldw = (n + 15) // 16 * 16
    V = np.full((ldw, w), np.nan, order="F"); Y = V.copy(order="F"); VT = V.copy(order="F"); tau = np.zeros(w)
    ret = lib.starneig_b200_panel(n, 0, n, w, A.ctypes.data, ld, V.ctypes.data, Y.ctypes.data, VT.ctypes.data, ldw, tau.ctypes.data)
    sn.starneig_node_finalize()
    for k in env: os.environ.pop(k)
    assert ret == 0
    return A[:n, :w].copy(), V[:n - 1].copy(), Y[:n - 1].copy(), VT[:n - 1].copy(), tau
ref = panel({"STARNEIG_B200_GEMV_LINEAR": "0"})
got = panel({})
assert all(np.isfinite(x).all() for x in got)
u = 2.0 ** -52
for a, b in zip(ref, got):
    assert np.abs(a - b).max() <= 200 * n * u * max(1.0, np.abs(a).max()), "linear GEMV differs from the sequential order"
    assert np.array_equal(a == 0.0, b == 0.0)
print("OK")
""" % (ROOT, SIM_LIB)


@pytest.mark.parametrize("sms,skew,seed", [(4, 4, 1), (6, 5, 2), (3, 3, 4)])
def test_sim_linear_gemv_with_lagging_blocks(simlib, sms, skew, seed):
    """One panel of a 600 x 600 matrix (three 256-row blocks of GEMV partials) while some blocks of the grid are scheduled far
    less often than the others and the thread schedule is shuffled: with GEMV linearity the GEMV warps of a column run
    concurrently with the look-ahead warps that derive its DLARFG scalars and s, so every ordering between the two must
    give the result of the sequential order (scalars first): V, Y, VT, tau and the panel columns within the tolerance."""
    r = subprocess.run(["python", "-c", _PANEL_CHILD, "600", "12"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUSIM_SMS=str(sms), CUSIM_SKEW=str(skew), CUSIM_SHUFFLE=str(seed)))
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.parametrize("gpus,n,pw,sms", [(1, 131, 24, 4), (1, 200, 40, 2), (2, 96, 16, 2), (1, 90, 16, 1), (2, 150, 24, 1)])
def test_sim_v_slab_in_shared_memory(sim, ora, gpus, n, pw, sms):
    """The CTA's rows of V kept in shared memory for the whole panel (FusedSmem, the default when they fit): the slab holds a
    copy of what goes to global memory, so H and Q are bitwise those of the kernel that re-reads global memory
    (STARNEIG_B200_FUSED_SLABS=0). One, two and three sub-tiles per CTA, one and two ranks."""
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=sms):
        with _Env(STARNEIG_B200_FUSED_SLABS=0):
            A0, Q0, st0 = _reduce(sim, ora, n, pw, gpus=gpus)
        assert st0["fused_slab_panels"][0] == st0["panels"] == st0["fused_panels"]
        A, Q, st = _reduce(sim, ora, n, pw, gpus=gpus, entrywise=False)          # the default
        assert st["fused_slab_panels"][1] == st["panels"], st["fused_slab_panels"]
        assert np.array_equal(A, A0) and np.array_equal(Q, Q0)


@pytest.mark.parametrize("n,pw,sms", [(131, 24, 4), (200, 40, 2), (90, 35, 1), (64, 64, 3), (47, 16, 4)])
def test_sim_q_backward_accumulation(sim, ora, n, pw, sms):
    """Q = I on entry (one GPU, full reduction): the reflectors of every panel are kept and Q = H_0 (H_1 (... H_K-1 I)) is
    formed after the last panel on the trailing blocks only (4/3 n^3 instead of 2 n^3 flops; engine.cuh, Rank::reduce).
    Same H bit for bit, Q equal to the forward product up to rounding, same invariants (checked by _reduce against the oracle,
    which accumulates forward like the reference)."""
    with _Env(CUSIM_SMS=sms, STARNEIG_B200_Q_BACKWARD=1):
        A, Q, st = _reduce(sim, ora, n, pw)
        assert st["q_backward"] == 1
        with _Env(STARNEIG_B200_Q_BACKWARD=0):
            A1, Q1, st1 = _reduce(sim, ora, n, pw)
        assert st1["q_backward"] == 0
        assert np.array_equal(A, A1)
        assert np.abs(Q[:n] - Q1[:n]).max() <= 50 * n * U
        if st["panels"] > 1:            # (one panel: both orders are the same product)
            assert not np.array_equal(Q, Q1) and st["gemm_flops"] < st1["gemm_flops"]


@pytest.mark.parametrize("gpus,n,pw,cb,sms", [(2, 96, 16, 8, 2), (3, 70, 16, 8, 2), (4, 150, 40, 16, 2), (8, 131, 24, 8, 1), (4, 20, 8, 16, 2)])
def test_sim_q_backward_accumulation_multi_rank(sim, ora, gpus, n, pw, cb, sms):
    """Several ranks: every rank forms its columns (block-cyclic like A) of Q = H_0 ... H_K-1 backward without any
    communication, then the ranks pull their row slabs out of the peers' column blocks (k_qcols_to_rows). The ranks agree on
    the order of accumulation by a sum over ranks. Same H bit for bit as the forward order, Q equal up to rounding."""
    with _Env(CUSIM_SMS=sms, STARNEIG_B200_Q_BACKWARD=1, STARNEIG_B200_COL_BLOCK=cb):
        A, Q, st = _reduce(sim, ora, n, pw, gpus=gpus)
        assert st["q_backward"] == 1
        with _Env(STARNEIG_B200_Q_BACKWARD=0):
            A1, Q1, st1 = _reduce(sim, ora, n, pw, gpus=gpus)
        assert st1["q_backward"] == 0
        assert np.array_equal(A, A1)
        assert np.abs(Q[:n] - Q1[:n]).max() <= 50 * n * U
        if st["panels"] > 1:
            assert not np.array_equal(Q, Q1)


def test_sim_q_backward_only_for_an_identity_q(sim, ora):
    """a general Q (the AED client's case), a partial reduction or too little device memory: forward, as before"""
    n, pw = 96, 16
    A0, Q0, ld = ora.fullpos(n, 2019)
    rng = np.random.default_rng(5)
    Qg, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Qin = np.zeros_like(Q0); Qin[:n, :n] = Qg
    with _Env(CUSIM_SMS=3, STARNEIG_B200_Q_BACKWARD=1, STARNEIG_B200_COL_BLOCK=8):
        # (general Q: Q H Q^T is not A0 any more, so the oracle comparison is done here instead of in _reduce)
        A, Q = A0.copy(order="F"), np.asfortranarray(Qin)
        sim.starneig_node_init(sim.STARNEIG_USE_ALL, 1, sim.STARNEIG_NO_MESSAGES)
        conf = sim.starneig_hessenberg_init_conf()
        conf.panel_width = pw
        assert sim.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, n, A, ld, Q, ld) == 0
        st = sim.get_stats()
        sim.starneig_node_finalize()
        assert st["q_backward"] == 0
        A2, Q2 = A0.copy(order="F"), np.asfortranarray(Qin)
        assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
        assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * np.abs(A2[:n]).max() and np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U
        Qneg = Q0.copy(order="F"); Qneg[3, 5] = -0.0                     # -0.0 is a zero
        _, _, st = _reduce(sim, ora, n, pw, given=(A0, Qneg, ld))
        assert st["q_backward"] == 1
        Qnan = Q0.copy(order="F"); Qnan[n - 1, 0] = 1e-300             # one tiny entry in a corner: not the identity
        _, _, st = _reduce(sim, ora, n, pw, given=(A0, Qnan, ld))
        assert st["q_backward"] == 0
        _, _, st = _reduce(sim, ora, n, pw, begin=0, end=n - 7, generator="partial")
        assert st["q_backward"] == 0
        _, _, st = _reduce(sim, ora, n, pw, gpus=2, given=(A0, Qnan, ld))        # only the LAST rank's slab differs from I
        assert st["q_backward"] == 0
        with _Env(CUSIM_FREE_MB=16):
            _, _, st = _reduce(sim, ora, n, pw)
        assert st["q_backward"] == 0


@pytest.mark.parametrize("gpus,n,pw,kc", [(1, 131, 24, 64), (2, 96, 16, 2048)])
def test_sim_gemv_staging_chunk(sim, ora, gpus, n, pw, kc):
    """STARNEIG_B200_GEMV_KC (columns of v staged per group at a time: 64 = many refills, 2048 = few): same sums in the same
    order => bitwise the same H and Q"""
    with _Env(STARNEIG_B200_COL_BLOCK=8, CUSIM_SMS=3):
        A, Q, _ = _reduce(sim, ora, n, pw, gpus=gpus)
        with _Env(STARNEIG_B200_GEMV_KC=kc):
            A1, Q1, _ = _reduce(sim, ora, n, pw, gpus=gpus)
    assert np.array_equal(A, A1) and np.array_equal(Q, Q1)




# ---------------------------------------------------------------------------------------------------------------------
# chain hand-off: the Hessenberg stage with H, Q left on the device, and the Reduce-shaped entry
# ---------------------------------------------------------------------------------------------------------------------
def _chain_callbacks(ora, n, log, device):
    """next stages of starneig_b200_SEP_SM_Reduce for the tests: `schur` = eigenvalues of H by LAPACK dhseqr (the oracle's
    stand-in for starneig_SEP_SM_Schur; H and Q are left as they are), `select` / `reorder_schur` record their calls.
    device = True: the Schur stage is handed the DEVICE pointers (host memory under the emulator)."""
    import ctypes
    from starneig_b200 import Chain, SCHUR_FN, SELECT_FN, REORDER_FN

    def schur(nn, pH, ldH, pQ, ldQ, preal, pimag):
        H = np.ctypeslib.as_array(ctypes.cast(pH, ctypes.POINTER(ctypes.c_double)), shape=(nn, ldH)).T      # (ldH x nn) view
        ev = ora.eigenvalues(nn, np.asfortranarray(H), ldH)
        np.ctypeslib.as_array(ctypes.cast(preal, ctypes.POINTER(ctypes.c_double)), shape=(nn,))[:] = ev.real
        np.ctypeslib.as_array(ctypes.cast(pimag, ctypes.POINTER(ctypes.c_double)), shape=(nn,))[:] = ev.imag
        log.append(("schur_device" if device else "schur", int(pH), ldH, int(pQ), ldQ))
        return 0

    def select(nn, pS, ldS, pred, arg, psel, pnum):
        sel = np.ctypeslib.as_array(ctypes.cast(psel, ctypes.POINTER(ctypes.c_int)), shape=(nn,))
        sel[:] = 0
        sel[: nn // 2] = 1
        pnum[0] = nn // 2
        log.append(("select", nn))
        return 0

    def reorder(nn, psel, pS, ldS, pQ, ldQ, preal, pimag):
        log.append(("reorder_schur", nn))
        return 0

    fns = (SCHUR_FN(schur), SELECT_FN(select), REORDER_FN(reorder))
    chain = Chain()
    if device:
        chain.schur_device = fns[0]
    else:
        chain.schur = fns[0]
    chain.select, chain.reorder_schur = fns[1], fns[2]
    return chain, fns        # keep the callbacks alive


def test_sim_hessenberg_stage_leaves_h_and_q_on_the_device(sim, ora):
    n = 90
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    sim.starneig_node_init(sim.STARNEIG_USE_ALL, 1, sim.STARNEIG_NO_MESSAGES)
    try:
        assert sim.starneig_SEP_SM_Hessenberg(n, A, ld, Q, ld) == 0
        A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
        ret, dH, lddH, dQ, lddQ = sim.hessenberg_stage(n, A1, ld, Q1, ld)
        assert ret == 0 and dH and dQ and lddH >= n and lddQ >= n
        assert np.array_equal(A1, A0) and np.array_equal(Q1, Q0)               # host arrays are inputs only
        assert sim.get_stats()["d2h_bytes"] == 0
        assert sim.stage_fetch(n, A1, ld, Q1, ld) == 0
        assert np.array_equal(A1, A) and np.array_equal(Q1, Q)                 # what the host-pointer call returns
        assert sim.stage_fetch(n + 1, A1, ld, Q1, ld) == -1
        # argument numbering of starneig_SEP_SM_Hessenberg
        assert sim.hessenberg_stage(0, A1, ld, Q1, ld)[0] == -1 and sim.hessenberg_stage(n, A1, n - 1, Q1, ld)[0] == -3
    finally:
        sim.starneig_node_finalize()
    sim.starneig_node_init(sim.STARNEIG_USE_ALL, 2, sim.STARNEIG_NO_MESSAGES)
    try:
        assert sim.hessenberg_stage(n, A1, ld, Q1, ld)[0] == sim.STARNEIG_GENERIC_ERROR        # one GPU only
    finally:
        sim.starneig_node_finalize()


@pytest.mark.parametrize("device", [False, True])
def test_sim_reduce_shaped_chain(sim, ora, device):
    """starneig_b200_SEP_SM_Reduce: Hessenberg here, Schur / Select / ReorderSchur from the caller (reference
    src/common/combined.c:45-98); eigenvalues of the chain against numpy's"""
    import ctypes
    from starneig_b200 import PREDICATE_FN
    n = 120
    A0, Q0, ld = ora.full(n, 12)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    real, imag = np.zeros(n), np.zeros(n)
    log = []
    chain, keep = _chain_callbacks(ora, n, log, device)
    sim.starneig_node_init(sim.STARNEIG_USE_ALL, 1, sim.STARNEIG_NO_MESSAGES)
    try:
        ret, num = sim.starneig_b200_SEP_SM_Reduce(n, A, ld, Q, ld, real, imag, chain)
        assert ret == 0 and [e[0] for e in log] == ["schur_device" if device else "schur"]
        ev = np.sort_complex(real + 1j * imag)
        want = np.sort_complex(np.linalg.eigvals(A0[:n]))
        d = np.abs(ev[:, None] - want[None, :]).min(axis=1)
        assert d.max() <= 1e-10 * np.linalg.norm(A0[:n])
        assert ora.hessenberg_form_violations(n, A, ld) == 0 and ora.residual_u(n, Q, ld, A, ld, A0, ld) <= 500
        # with a predicate the selection stages run too
        A, Q = A0.copy(order="F"), Q0.copy(order="F")
        selected = np.zeros(n, dtype=np.int32)
        pred = PREDICATE_FN(lambda re, im, arg: int(re < 0.0))
        ret, num = sim.starneig_b200_SEP_SM_Reduce(n, A, ld, Q, ld, real, imag, chain, predicate=pred, selected=selected)
        assert ret == 0 and num == n // 2 and [e[0] for e in log][-2:] == ["select", "reorder_schur"] and selected.sum() == n // 2
        # an incomplete chain is an argument error (the reference's numbering stops at 11)
        from starneig_b200 import Chain
        assert sim.starneig_b200_SEP_SM_Reduce(n, A, ld, Q, ld, real, imag, Chain())[0] == -12
    finally:
        sim.starneig_node_finalize()
