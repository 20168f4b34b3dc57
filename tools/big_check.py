"""Largest configuration on ONE GPU, device-resident (BASELINE.json configs[3] is n = 50000: A and Q are 20 GB each and fit
one B200's 180 GB): reduction through starneig_b200_hessenberg_device, then the reference driver's invariants evaluated
on the GPU with torch FP64 matmuls (the CPU oracle cannot reach this size): exact-zero Hessenberg form, residual
|Q H Q^T - A|_F / |A|_F, orthogonality |Q Q^T - I|_F / sqrt(n), trace. usage: big_check.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starneig_b200 as sn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ld = (n + 15) // 16 * 16
u = 2.0 ** -52
gen = torch.Generator(device="cuda").manual_seed(2019)
A0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)      # row c = column c of the matrix
A = A0.clone()
Q = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
Q.diagonal().fill_(1.0)
sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
sn.set_profile_level(1)
torch.cuda.synchronize()
t0 = time.time()
ret = sn.hessenberg_device(n, A, ld, Q, ld)
st = sn.get_stats()
print(f"n {n} panel_width {st['panel_width']} panels {st['panels']} ret {ret} wall {time.time() - t0:.1f}s device_ms {st['device_ms']:.0f} "
      f"GFLOP/s {10 / 3 * n ** 3 / st['device_ms'] / 1e6:.0f} col {st['panel_ms']:.0f} trail {st['trail_ms']:.0f} deferred {st['other_ms']:.0f} "
      f"gemv_ms {st['gemv_ms']:.0f} ({st['gemv_timed_bytes'] / max(st['gemv_ms'], 1e-9) / 1e6:.0f} GB/s) launches {st['kernel_launches']}", flush=True)
sn.starneig_node_finalize()
assert ret == 0
# matrices as torch sees them: M_t[c, r] = M(r, c), i.e. the transpose. H^T must be lower Hessenberg in torch's view.
Ht, Qt, A0t = A[:, :n], Q[:, :n], A0[:, :n]
bad = 0
for c0 in range(0, n, 4096):
    blk = Ht[c0:c0 + 4096]                     # columns c0.. of H; entries (r, c) with r > c + 1 must be exactly zero
    rows = torch.arange(n, device="cuda")[None, :]
    cols = torch.arange(c0, min(n, c0 + 4096), device="cuda")[:, None]
    bad += int(((rows > cols + 1) & (blk != 0.0)).sum())
finite = bool(torch.isfinite(Ht).all()) and bool(torch.isfinite(Qt).all())
tr = abs(float(Ht.diagonal().sum() - A0t.diagonal().sum())) / abs(float(A0t.diagonal().sum()))
normA = float(torch.linalg.norm(A0t))
# (Q H Q^T)^T = Q_t^T H_t Q_t in torch's view
T = Qt.T @ Ht
R = T @ Qt
del T
R -= A0t
res = float(torch.linalg.norm(R)) / normA / u
del R
G = Qt.T @ Qt                                   # (Q Q^T)^T
G.diagonal().sub_(1.0)
orth = float(torch.linalg.norm(G)) / n ** 0.5 / u
print(f"finite {finite} form_violations {bad} residual {res:.1f} u orthogonality {orth:.1f} u trace_rel_err {tr:.2e} "
      f"(bounds: 10 n u = {10 * n} u, reference driver warn 500 u / fail 10000 u)")
assert finite and bad == 0 and res < 500 and orth < 500
print("OK")
