"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when built, to the reference's own
Hessenberg sources (oracle/_ref/libstarneig_ref.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by starneig_b200.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libstarneig_ref.so")

_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous
    return a.ctypes.data_as(_dp)


def build(quiet=True):
    """(Re)build the oracle; `_ref` is only rebuilt where /root/reference exists."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Conf(ctypes.Structure):
    _fields_ = [("tile_size", ctypes.c_int), ("panel_width", ctypes.c_int)]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = lib = ctypes.CDLL(ORACLE_SO)
        lib.oracle_residual_u.restype = ctypes.c_double
        lib.oracle_orthogonality_u.restype = ctypes.c_double
        lib.oracle_frobenius.restype = ctypes.c_double
        lib.oracle_check_hessenberg_form.restype = ctypes.c_long
        lib.oracle_wall_seconds.restype = ctypes.c_double
        lib.oracle_prand_init.argtypes = [ctypes.c_uint]

    # ---- generators (reference test driver) ----
    @staticmethod
    def ld_for(n):
        """leading dimension used by the reference test driver: rows rounded up to 8 doubles
        (test/common/common.c:96-112)"""
        return (n + 7) // 8 * 8

    def prand_init(self, seed):
        self.lib.oracle_prand_init(seed)

    def prand(self):
        return self.lib.oracle_prand()

    def alloc(self, n, ld=None):
        ld = ld or self.ld_for(n)
        return np.zeros((ld, n), dtype=np.float64, order="F"), ld

    def fullpos(self, n, seed=2019, ld=None):
        """A = fullpos LCG matrix, Q = I (test/hessenberg/experiment.c:85-113)"""
        A, ld = self.alloc(n, ld)
        Q, _ = self.alloc(n, ld)
        self.prand_init(seed)
        self.lib.oracle_fill_fullpos(n, _p(A), ld)
        self.lib.oracle_fill_identity(n, _p(Q), ld)
        return A, Q, ld

    def full(self, n, seed=2019, ld=None):
        A, ld = self.alloc(n, ld)
        Q, _ = self.alloc(n, ld)
        self.prand_init(seed)
        self.lib.oracle_fill_full(n, _p(A), ld)
        self.lib.oracle_fill_identity(n, _p(Q), ld)
        return A, Q, ld

    def partial(self, n, begin, end, seed=2019, ld=None):
        """upper triangular + full block [begin,end) (test/misc/partial_hessenberg.c:142-157)"""
        A, ld = self.alloc(n, ld)
        Q, _ = self.alloc(n, ld)
        self.prand_init(seed)
        self.lib.oracle_fill_partial(n, begin, end, _p(A), ld)
        self.lib.oracle_fill_identity(n, _p(Q), ld)
        return A, Q, ld

    # ---- solvers ----
    def set_threads(self, t):
        self.lib.oracle_set_threads(t)

    def hessenberg_port(self, n, A, ldA, Q, ldQ, begin=0, end=None, panel_width=-1):
        end = n if end is None else end
        return self.lib.oracle_hessenberg_port(n, begin, end, panel_width, _p(A), ldA, _p(Q), ldQ)

    def hessenberg_lapack(self, n, A, ldA, Q, ldQ):
        return self.lib.oracle_hessenberg_lapack(n, _p(A), ldA, _p(Q), ldQ)

    def default_panel_width(self, n):
        return self.lib.oracle_default_panel_width(n)

    # ---- checks (reference test driver) ----
    def hessenberg_form_violations(self, n, A, ldA, begin=0, end=None, check_outside=False):
        end = n if end is None else end
        return self.lib.oracle_check_hessenberg_form(n, begin, end, int(check_outside), _p(A), ldA)

    def residual_u(self, n, Q, ldQ, H, ldH, A0, ldA0):
        return self.lib.oracle_residual_u(n, _p(Q), ldQ, _p(H), ldH, _p(A0), ldA0)

    def orthogonality_u(self, n, Q, ldQ):
        return self.lib.oracle_orthogonality_u(n, _p(Q), ldQ)

    def eigenvalues(self, n, H, ldH):
        wr = np.zeros(n); wi = np.zeros(n)
        info = self.lib.oracle_hessenberg_eigenvalues(n, _p(H), ldH, wr.ctypes.data_as(_dp), wi.ctypes.data_as(_dp))
        assert info == 0, f"dhseqr failed: {info}"
        return wr + 1j * wi


class Reference:
    """The reference's own src/hessenberg + src/common sources, compiled against the StarPU stand-in
    (oracle/ref_shim). Only available where oracle/_ref has been built."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = ctypes.CDLL(REF_SO)

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def set_threads(self, t):
        self.lib.oracle_ref_set_threads(t)

    def set_workers(self, w):
        """number of workers REPORTED to the reference (its default tile size depends on it)"""
        self.lib.oracle_ref_set_workers(w)

    def set_executors(self, count):
        """threads that execute the task graph in parallel under its data dependencies (0: every task inline at its
        insertion point, the default and what the fixtures use); oracle/ref_shim/mini_starpu.c"""
        self.lib.oracle_ref_set_executors(count)

    def tasks_executed(self, reset=False):
        self.lib.oracle_ref_tasks_executed.restype = ctypes.c_ulong
        return int(self.lib.oracle_ref_tasks_executed(1 if reset else 0))

    def hessenberg(self, n, A, ldA, Q, ldQ):
        return self.lib.starneig_SEP_SM_Hessenberg(n, _p(A), ldA, _p(Q), ldQ)

    def hessenberg_expert(self, n, A, ldA, Q, ldQ, begin=0, end=None, tile_size=-1, panel_width=-1):
        end = n if end is None else end
        conf = Conf(tile_size, panel_width)
        return self.lib.starneig_SEP_SM_Hessenberg_expert(ctypes.byref(conf), n, begin, end, _p(A), ldA, _p(Q), ldQ)
