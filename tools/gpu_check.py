"""Development check run on the GPU box: every kernel against numpy / the CPU oracle, with diagnostics.
(The pytest suite under tests/ is the gate; this script prints more detail while developing.)"""
import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starneig_b200 as sn
from oracle.oracle import Oracle

ora = Oracle()
L = sn.lib()
sn.starneig_node_init(-1, 1, sn.STARNEIG_NO_MESSAGES)
dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
ok_all = True

def colmajor(m, n, ld=None, pad=0):
    """device column-major m x n with leading dimension ld: torch tensor of shape (n, ld) == F-order (ld, n)"""
    ld = ld or m
    host = rng.standard_normal((ld, n))
    host = np.asfortranarray(host)
    t = torch.from_numpy(np.ascontiguousarray(host.T)).to(dev)   # t[j, i] = host[i, j]; memory = column-major
    return host, t

def check(name, got, ref, tol):
    global ok_all
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    flag = "ok " if err <= tol else "BAD"
    if err > tol or not np.isfinite(err): ok_all = False
    print(f"[{flag}] {name}: rel err {err:.3e}")

def test_gemm(ta, tb, m, n, k, off=0):
    # op(A) m x k, op(B) k x n
    Ar, Ac = (m, k) if ta == 'N' else (k, m)
    Br, Bc = (k, n) if tb == 'N' else (n, k)
    lda, ldb, ldc = Ar + off + 3, Br + off + 5, m + off + 1
    Ah, At = colmajor(Ar + off, Ac, lda); Bh, Bt = colmajor(Br + off, Bc, ldb); Ch, Ct = colmajor(m + off, n, ldc)
    alpha, beta = (-1.0, 1.0) if (ta, tb) == ('N', 'T') else (1.0, 0.0)
    opA = Ah[off:off+Ar, :] if ta == 'N' else Ah[off:off+Ar, :].T
    opB = Bh[off:off+Br, :] if tb == 'N' else Bh[off:off+Br, :].T
    ref = alpha * opA @ opB + beta * Ch[off:off+m, :]
    r = L.starneig_b200_dgemm(ta.encode(), tb.encode(), m, n, k, alpha, At.data_ptr() + 8*off, lda, Bt.data_ptr() + 8*off, ldb,
                              beta, Ct.data_ptr() + 8*off, ldc)
    got = Ct.cpu().numpy().T[off:off+m, :]
    untouched = np.array_equal(Ct.cpu().numpy().T[:off, :], Ch[:off, :]) and np.array_equal(Ct.cpu().numpy().T[off+m:, :], Ch[off+m:, :])
    check(f"dgemm {ta}{tb} m={m} n={n} k={k} off={off} ret={r} untouched={untouched}", got, ref, 1e-12)

def test_gemv(m, k, off=0):
    lda = (m + off + 9) // 2 * 2
    Ah, At = colmajor(m + off, k, lda)
    v = rng.standard_normal(k); v[0] = 1.0
    vt = torch.from_numpy(v).to(dev); yt = torch.zeros(m, dtype=torch.float64, device=dev)
    ms = ctypes.c_float(0)
    r = L.starneig_b200_gemv(m, k, At.data_ptr() + 8*off, lda, vt.data_ptr(), yt.data_ptr(), 3, ctypes.byref(ms))
    check(f"gemv m={m} k={k} off={off} ret={r} {ms.value*1e3:.1f} us {m*k*8/ms.value/1e6:.0f} GB/s", yt.cpu().numpy(), Ah[off:off+m, :] @ v, 1e-12)

def test_hess(n, pw, begin=0, end=None, gen="fullpos", ld=None, tol=1e-11, compare=True):
    global ok_all
    end = n if end is None else end
    if gen == "fullpos": A0, Q0, ld = ora.fullpos(n, ld=ld)
    elif gen == "partial": A0, Q0, ld = ora.partial(n, begin, end, ld=ld)
    else: A0, Q0, ld = ora.full(n, ld=ld)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    conf = sn.starneig_hessenberg_init_conf(); conf.panel_width = pw
    t = time.time()
    r = sn.starneig_SEP_SM_Hessenberg_expert(conf, n, begin, end, A, ld, Q, ld)
    dt = time.time() - t
    st = sn.get_stats()
    form = ora.hessenberg_form_violations(n, A, ld, begin, end, check_outside=(gen == "partial"))
    res = ora.residual_u(n, Q, ld, A, ld, A0, ld) if n <= 6000 else float("nan")
    orth = ora.orthogonality_u(n, Q, ld) if n <= 6000 else float("nan")
    if st['gemv_timed_launches'] > 0:
        k = st['gemv_timed_launches']
        print(f"      per timed column: finish_update {1e3*st['finish_update_ms']/k:.1f} us, reflector {1e3*st['reflector_ms']/k:.1f} us, gemv {1e3*st['gemv_ms']/k:.1f} us ({st['gemv_timed_bytes']/max(st['gemv_ms'],1e-9)/1e6:.0f} GB/s)")
    msg = f"hess n={n} pw={pw} [{begin},{end}) {gen} ret={r} wall={dt*1e3:.1f} ms dev={st['device_ms']:.1f} ms (panel {st['panel_ms']:.1f} trail {st['trail_ms']:.1f} other {st['other_ms']:.1f}) form={form} res={res:.1f}u orth={orth:.1f}u"
    good = r == 0 and form == 0 and (not np.isfinite(res) or res < 1000) and (not np.isfinite(orth) or orth < 1000) and np.isfinite(A[:n]).all()
    if compare:
        A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
        ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end, pw)
        eh = np.abs(A[:n] - A2[:n]).max() / np.abs(A2[:n]).max(); eq = np.abs(Q[:n] - Q2[:n]).max()
        msg += f" |H-Hport|/max|H|={eh:.2e} |Q-Qport|={eq:.2e}"
        good = good and eh < tol and eq < tol
    gf = 10.0 / 3.0 * n**3 / (st['device_ms'] * 1e-3) / 1e9 if st['device_ms'] > 0 else 0
    print(f"[{'ok ' if good else 'BAD'}] {msg} | {gf:.0f} GFLOP/s(dev)")
    if not good: ok_all = False

which = sys.argv[1:] or ["gemm", "gemv", "hess_small", "hess_mid"]
if "gemm" in which:
    for (ta, tb) in [("N", "T"), ("T", "N"), ("N", "N")]:
        for (m, n, k, off) in [(128, 128, 16, 0), (64, 40, 8, 0), (257, 131, 37, 1), (500, 312, 1000, 3), (1000, 96, 300, 2), (37, 5, 3, 1), (2000, 280, 2100, 1)]:
            test_gemm(ta, tb, m, n, k, off)
if "gemv" in which:
    for (m, k, off) in [(256, 16, 0), (300, 300, 1), (1000, 777, 2), (1, 1, 0), (5, 9, 1), (4097, 4000, 3), (16384, 16384, 1)]:
        test_gemv(m, k, off)
if "gemm_perf" in which:
    for (ta, tb, m, n, k) in [("N","T",16000,15700,312), ("T","N",15700,312,16000), ("N","N",16000,312,16000), ("N","T",4000,3700,312), ("T","N",3700,312,4000), ("N","N",16000,312,4000), ("N","T",16000,16000,296)]:
        Ar, Ac = (m, k) if ta == 'N' else (k, m)
        Br, Bc = (k, n) if tb == 'N' else (n, k)
        At = torch.rand((Ac, Ar), dtype=torch.float64, device=dev); Bt = torch.rand((Bc, Br), dtype=torch.float64, device=dev); Ct = torch.rand((n, m), dtype=torch.float64, device=dev)
        alpha, beta = (-1.0, 1.0) if (ta, tb) == ('N', 'T') else (1.0, 0.0)
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            L.starneig_b200_dgemm(ta.encode(), tb.encode(), m, n, k, alpha, At.data_ptr(), Ar, Bt.data_ptr(), Br, beta, Ct.data_ptr(), m)
            dt = time.perf_counter() - t0
        print(f"gemm_perf {ta}{tb} m={m} n={n} k={k}: {dt*1e3:.3f} ms {2.0*m*n*k/dt/1e12:.2f} TFLOP/s")
if "hess_small" in which:
    for (n, pw) in [(1, 8), (2, 8), (3, 8), (9, 8), (10, 8), (17, 8), (40, 8), (47, 16), (88, 35), (100, 100), (333, 45), (554, 170)]:
        test_hess(n, pw)
    test_hess(88, 16, 22, 66, gen="partial")
    test_hess(333, 35, 83, 249, gen="partial")
    test_hess(201, 32, gen="full", ld=230)
if "hess_mid" in which:
    sn.set_profile_level(2)
    test_hess(1000, -1)
    test_hess(2000, -1)
    test_hess(2000, -1)
    test_hess(4000, -1, compare=False)
if "hess_big" in which:
    sn.set_profile_level(2)
    test_hess(10000, -1, compare=False)
    st = sn.get_stats(); print(st)
    print(f"gemv (sampled): {st['gemv_ms']:.1f} ms for {st['gemv_timed_bytes']/1e9:.1f} GB -> {st['gemv_timed_bytes']/st['gemv_ms']/1e6:.0f} GB/s")
if "hess_20k" in which:
    sn.set_profile_level(2)
    test_hess(20000, -1, compare=False)
    st = sn.get_stats(); print(st)
    print(f"gemv (sampled): {st['gemv_ms']:.1f} ms for {st['gemv_timed_bytes']/1e9:.1f} GB -> {st['gemv_timed_bytes']/st['gemv_ms']/1e6:.0f} GB/s")
sn.starneig_node_finalize()
print("ALL OK" if ok_all else "SOME FAILED")
