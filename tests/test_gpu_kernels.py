"""GPU parity tests of the individual kernels, called through the C ABI (include/starneig_b200.h) and
checked against float64 numpy on the same seeded inputs."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _colmajor(rng, rows, cols, ld):
    import torch
    host = np.asfortranarray(rng.standard_normal((ld, cols)))
    dev = torch.from_numpy(np.ascontiguousarray(host.T)).cuda()     # same bytes as column-major (ld x cols)
    return host, dev


def _back(dev):
    return dev.cpu().numpy().T


# NT: rank-k update (cpu.c:315,433,552); TN: W = A^T V (cpu.c:373); NN: W = A V (cpu.c:492)
@pytest.mark.parametrize("ta,tb", [("N", "T"), ("T", "N"), ("N", "N")])
@pytest.mark.parametrize("m,n,k,off", [(128, 128, 16, 0), (1, 1, 1, 0), (37, 5, 3, 1), (257, 131, 37, 1),
                                       (500, 312, 1000, 3), (1000, 96, 300, 2), (3000, 280, 2999, 1),
                                       (129, 105, 17, 0), (700, 296, 5000, 5)])
def test_dgemm(node, ta, tb, m, n, k, off):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    L = node.lib()
    Ar, Ac = (m, k) if ta == "N" else (k, m)
    Br, Bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = Ar + off + 3, Br + off + 5, m + off + 1
    Ah, Ad = _colmajor(rng, Ar + off, Ac, lda)
    Bh, Bd = _colmajor(rng, Br + off, Bc, ldb)
    Ch, Cd = _colmajor(rng, m + off, n, ldc)
    alpha, beta = (-1.0, 1.0) if (ta, tb) == ("N", "T") else (1.0, 0.0)
    opA = Ah[off:off + Ar] if ta == "N" else Ah[off:off + Ar].T
    opB = Bh[off:off + Br] if tb == "N" else Bh[off:off + Br].T
    want = alpha * opA @ opB + beta * Ch[off:off + m]
    ret = L.starneig_b200_dgemm(ta.encode(), tb.encode(), m, n, k, alpha, Ad.data_ptr() + 8 * off, lda,
                                Bd.data_ptr() + 8 * off, ldb, beta, Cd.data_ptr() + 8 * off, ldc)
    assert ret == 0
    got = _back(Cd)
    # FP64 accumulation in a different order than numpy: tolerance k * eps * |A||B|
    bound = 4 * k * 2.0 ** -52 * (np.abs(opA) @ np.abs(opB)).max() + 1e-300
    assert np.abs(got[off:off + m] - want).max() <= bound
    # rows outside the window are untouched (the kernels write nothing beyond M x N)
    assert np.array_equal(got[:off], Ch[:off]) and np.array_equal(got[off + m:], Ch[off + m:])


def test_dgemm_rejects_unsupported(node):
    import torch
    x = torch.zeros(64, dtype=torch.float64, device="cuda")
    L = node.lib()
    assert L.starneig_b200_dgemm(b"T", b"T", 4, 4, 4, 1.0, x.data_ptr(), 4, x.data_ptr(), 4, 0.0, x.data_ptr(), 4) == 4


# y = A v: the compute_column codelet (cpu.c:163-224, cuda.cu:62-150)
@pytest.mark.parametrize("m,k,off", [(1, 1, 0), (5, 9, 1), (256, 16, 0), (300, 300, 1), (1000, 777, 2),
                                     (4097, 4000, 3), (2, 3000, 1), (9000, 33, 0)])
def test_gemv(node, m, k, off):
    import torch
    rng = np.random.default_rng(m + 13 * k)
    L = node.lib()
    lda = (m + off + 9) // 2 * 2
    Ah, Ad = _colmajor(rng, m + off, k, lda)
    v = rng.standard_normal(k); v[0] = 1.0
    vd = torch.from_numpy(v).cuda()
    yd = torch.full((m,), np.nan, dtype=torch.float64, device="cuda")
    ms = ctypes.c_float(0)
    assert L.starneig_b200_gemv(m, k, Ad.data_ptr() + 8 * off, lda, vd.data_ptr(), yd.data_ptr(), 1, ctypes.byref(ms)) == 0
    want = Ah[off:off + m] @ v
    bound = 4 * k * 2.0 ** -52 * (np.abs(Ah[off:off + m]) @ np.abs(v)).max() + 1e-300
    assert np.abs(yd.cpu().numpy() - want).max() <= bound


def test_gemv_is_deterministic(node):
    import torch
    rng = np.random.default_rng(3)
    L = node.lib()
    m = k = 3000
    Ah, Ad = _colmajor(rng, m, k, m)
    v = rng.standard_normal(k); v[0] = 1.0
    vd = torch.from_numpy(v).cuda()
    outs = []
    for _ in range(3):
        yd = torch.zeros(m, dtype=torch.float64, device="cuda")
        L.starneig_b200_gemv(m, k, Ad.data_ptr(), m, vd.data_ptr(), yd.data_ptr(), 1, None)
        outs.append(yd.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


# one panel: prepare_column / compute_column / finish_column (cpu.c:50-285) against a numpy restatement
@pytest.mark.parametrize("n,i,w", [(64, 0, 8), (200, 0, 32), (300, 40, 45), (700, 101, 96), (513, 0, 280)])
def test_panel_against_numpy(node, n, i, w):
    import torch
    rng = np.random.default_rng(n + i)
    L = node.lib()
    ld = (n + 15) // 16 * 16
    A0 = np.asfortranarray(rng.standard_normal((ld, n)))
    Ad = torch.from_numpy(np.ascontiguousarray(A0.T)).cuda()
    m = n - i - 1
    ldw = (m + 63) // 64 * 64
    Vd = torch.zeros((w, ldw), dtype=torch.float64, device="cuda")
    Yd = torch.zeros((w, ldw), dtype=torch.float64, device="cuda")
    VTd = torch.zeros((w, ldw), dtype=torch.float64, device="cuda")
    tau = np.zeros(w)
    assert L.starneig_b200_panel(n, i, n, w, Ad.data_ptr(), ld, Vd.data_ptr(), Yd.data_ptr(), VTd.data_ptr(), ldw,
                                 tau.ctypes.data) == 0
    A1 = _back(Ad)[:n]; V = _back(Vd)[:m]; Y = _back(Yd)[:m]; VT = _back(VTd)[:m]

    # dense restatement in numpy
    A = A0[:n].copy()
    Vr = np.zeros((m, w)); Yr = np.zeros((m, w)); T = np.zeros((w, w))
    for j in range(w):
        c = i + j
        p = A[i + 1:, c].copy()
        if j > 0:
            p -= Yr[:, :j] @ Vr[j - 1, :j]
            p -= Vr[:, :j] @ (T[:j, :j].T @ (Vr[:, :j].T @ p))
        alpha, x = p[j], p[j + 1:]
        xn = np.linalg.norm(x)
        if len(x) == 0 or xn == 0:
            t, beta, v = 0.0, alpha, np.zeros_like(x)
        else:
            beta = -np.copysign(np.hypot(alpha, xn), alpha)
            t = (beta - alpha) / beta
            v = x / (alpha - beta)
        Vr[j, j] = 1.0; Vr[j + 1:, j] = v
        A[i + 1:i + 1 + j, c] = p[:j]; A[i + 1 + j, c] = beta; A[i + 2 + j:, c] = 0.0
        y = A[i + 1:, c + 1:] @ Vr[j:, j]
        s = Vr[j:, :j].T @ Vr[j:, j]
        Yr[:, j] = t * (y - Yr[:, :j] @ s)
        T[:j, j] = -t * (T[:j, :j] @ s); T[j, j] = t
    scale = max(1.0, np.abs(A).max())
    tol = 500 * n * 2.0 ** -52
    assert np.abs(tau - np.diag(T)).max() <= tol
    assert np.abs(A1 - A).max() <= tol * scale
    assert np.abs(V - Vr).max() <= tol
    assert np.abs(Y - Yr).max() <= tol * scale
    assert np.abs(VT - Vr @ T).max() <= tol                      # VT == V*T, the compact-WY factor applied
    # exact zeros below the sub-diagonal of the reduced columns, exact unit diagonal / zeros in V
    for j in range(w):
        assert np.all(A1[i + 2 + j:, i + j] == 0.0)
        assert V[j, j] == 1.0 and np.all(V[:j, j] == 0.0)
