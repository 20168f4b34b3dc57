import os
import sys

import pytest

# Virtual ranks (several ranks sharing one device in tests/test_gpu_multi.py) need their streams on distinct hardware
# queues, or a spinning consumer kernel could sit in front of its producer. Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """the reference's own sources built against the StarPU stand-in, inline schedule (absent => skip)"""
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = Reference()
    r.set_threads(1)
    return r


@pytest.fixture(scope="session")
def sn():
    import starneig_b200
    return starneig_b200


@pytest.fixture()
def node(sn):
    """initialised node with one GPU, finalised afterwards"""
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    yield sn
    sn.starneig_node_finalize()


def same_zero_pattern(A, Aref, n, scale=None):
    """The exact zeros of two reductions of the same matrix agree: everything strictly below the first sub-diagonal is a
    STRUCTURAL zero (written as 0.0 by both: reference src/hessenberg/cpu.c:153) and must match exactly; on and above the
    sub-diagonal an entry that is zero only through cancellation (tau = 0 inputs, the last sub-diagonal entry of such a
    matrix) may come out as 0.0 in one and as rounding noise in the other, so there a mismatch is accepted if both values are
    below 200 n u max|Aref|."""
    import numpy as np
    u = 2.0 ** -52
    Z, Zr = A[:n] == 0.0, Aref[:n] == 0.0
    low = np.tril(np.ones((n, A.shape[1]), dtype=bool), -2)
    if not np.array_equal(Z & low, Zr & low):
        return False
    diff = (Z != Zr) & ~low
    if not diff.any():
        return True
    tol = 200 * n * u * max(1.0, float(np.abs(Aref[:n]).max()) if scale is None else scale)
    return bool((np.maximum(np.abs(A[:n]), np.abs(Aref[:n]))[diff] <= tol).all())


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("prand", "chain")))


# Structured inputs in which DLARFG meets x = 0 (tau = 0, H = I: reference src/hessenberg/cpu.c:140, LAPACK dlarfg's early
# return) in some or all columns, plus the shape the Schur stage's AED step hands to the Hessenberg reduction
# (src/schur/core.c:893-929: a deflation window = upper triangular + a spike in its first column). `entrywise` says whether
# H and Q may be compared entry by entry with another backward-stable reduction: the spike window has exactly zero
# sub-diagonal entries in H, so H and Q are not uniquely determined to working precision (the reference port and LAPACK
# differ by 1e7 * n * u on it) and only the invariants are checked.
STRUCTURED = ["zero", "identity", "upper_triangular", "already_hessenberg", "zero_columns", "block_triangular", "aed_spike"]


def structured_input(ora, name, n, seed=7):
    import numpy as np
    A0, Q0, ld = ora.full(n, seed)
    A = A0.copy(order="F")
    entrywise = True
    if name == "zero":
        A[:n] = 0.0
    elif name == "identity":
        A[:n] = np.eye(n)
    elif name == "upper_triangular":
        A[:n] = np.triu(A0[:n])
    elif name == "already_hessenberg":
        A[:n] = np.triu(A0[:n], -1)
    elif name == "zero_columns":
        A[:n, min(3, n - 1)] = 0.0
        A[:n, n // 2] = 0.0
        A[min(5, n):n, min(10, n - 1)] = 0.0
    elif name == "block_triangular":
        A[n // 2:n, :n // 2] = 0.0
    elif name == "aed_spike":
        A[:n] = np.triu(A0[:n])
        A[1:3 * n // 4, 0] = A0[1:3 * n // 4, 1]
        entrywise = False
    else:
        raise ValueError(name)
    return A, Q0, ld, entrywise


def aed_window_check(sn, ora, n, end, pw, gpus=None):
    """The Schur stage's AED client (reference src/schur/core.c:893-929): a deflation window (upper triangular + spike,
    only [0, end) reduced) whose transformations are accumulated into a non-identity local Q. Invariants only (see
    STRUCTURED): exact Hessenberg form inside the block, untouched zeros below it, Q orthogonal, Q H Q^T = Q0 A0 Q0^T,
    and the same exact-zero pattern as the reference port. `gpus`: initialise / finalise the node here (emulator tests)."""
    import numpy as np
    u = 2.0 ** -52
    A0, _, ld, _ = structured_input(ora, "aed_spike", n)
    A0[end:n, :end] = 0.0
    Qr, _ = np.linalg.qr(np.random.default_rng(3).standard_normal((n, n)))
    Q0 = np.zeros((ld, n), order="F")
    Q0[:n] = Qr
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    if gpus is not None:
        sn.starneig_node_init(sn.STARNEIG_USE_ALL, gpus, sn.STARNEIG_NO_MESSAGES)
    try:
        conf = sn.starneig_hessenberg_init_conf()
        conf.panel_width = pw
        assert sn.starneig_SEP_SM_Hessenberg_expert(conf, n, 0, end, A, ld, Q, ld) == 0
    finally:
        if gpus is not None:
            sn.starneig_node_finalize()
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, end, pw) == 0
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert np.count_nonzero(np.tril(A[:end, :end], -2)) == 0 and np.count_nonzero(A[end:n, :end]) == 0
    assert same_zero_pattern(A, A2, n)
    assert ora.orthogonality_u(n, Q, ld) <= 500
    res = np.linalg.norm(Q[:n] @ A[:n] @ Q[:n].T - Qr @ A0[:n] @ Qr.T) / np.linalg.norm(A0[:n]) / u
    assert res <= 500, res
    assert np.array_equal(A[n:], A0[n:]) and np.array_equal(Q[n:], Q0[n:])      # padding rows untouched
