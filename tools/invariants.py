"""The reference test driver's acceptance checks for a Hessenberg reduction, evaluated on a GPU with torch FP64 matmuls
(checker code, never on the product path; the CPU oracle's versions of the same checks are too slow beyond n ~ 4000):

  * form          entries (r, c) with r > c + 1 inside [begin, end) are EXACTLY zero
                  (reference test/common/hooks.c:434-487, the `hessenberg` hook)
  * residual      |Q H Q^T - A|_F / |A|_F   in units of u = 2^-52     (test/common/checks.c:180-208, hooks.c:258-353)
  * orthogonality |Q Q^T - I|_F / sqrt(n)   in units of u
  * trace         |tr H - tr A| / |tr A|

Storage convention of the callers (bench.py, tools/big_check*.py, tests): a torch tensor M_t of shape (n, ld) holds the
column-major matrix M with M_t[c, r] = M(r, c), i.e. torch sees the transpose.
Thresholds: the reference driver warns above 500 u and fails above 10000 u; BASELINE.json asks for <= 10 n u."""
import torch

U = 2.0 ** -52
WARN_U = 500.0


def form_violations(Ht, n, begin=0, end=None, block=4096):
    """entries below the first sub-diagonal of columns [begin, end) that are not exactly zero (rows < end)"""
    end = n if end is None else end
    dev = Ht.device
    bad = 0
    rows = torch.arange(n, device=dev)[None, :]
    for c0 in range(begin, end, block):
        c1 = min(end, c0 + block)
        cols = torch.arange(c0, c1, device=dev)[:, None]
        blk = Ht[c0:c1, :n]
        bad += int(((rows > cols + 1) & (rows < end) & (blk != 0.0)).sum())
    return bad


def evaluate(A0t, Ht, Qt, n, begin=0, end=None, Q0t=None):
    """A0t, Ht, Qt: (n, >= n) tensors on one CUDA device, transposed storage (see above). Q0t: the initial Q if it was not
    the identity (the reduction then satisfies Q H Q^T = Q0 A0 Q0^T). Returns a dict of plain floats / ints."""
    A0, H, Q = A0t[:, :n], Ht[:, :n], Qt[:, :n]
    finite = bool(torch.isfinite(H).all()) and bool(torch.isfinite(Q).all())
    bad = form_violations(Ht, n, begin, end)
    trA = float(A0.diagonal().sum())
    tr = abs(float(H.diagonal().sum()) - trA) / max(abs(trA), 1e-300)
    normA = float(torch.linalg.norm(A0))
    # (Q H Q^T)^T = Q_t^T H_t Q_t in torch's view
    T = Q.T @ H
    R = T @ Q
    del T
    if Q0t is None:
        R -= A0
    else:
        Q0 = Q0t[:, :n]
        T = Q0.T @ A0
        R -= T @ Q0
        del T
    res = float(torch.linalg.norm(R)) / max(normA, 1e-300) / U
    del R
    G = Q.T @ Q                                   # (Q Q^T)^T
    G.diagonal().sub_(1.0)
    orth = float(torch.linalg.norm(G)) / n ** 0.5 / U
    del G
    return {"n": n, "finite": finite, "form_violations": bad, "residual_u": res, "orthogonality_u": orth,
            "trace_rel_err": tr, "bound_u": min(WARN_U, 10.0 * n),
            "ok": bool(finite and bad == 0 and res <= min(WARN_U, 10.0 * n) and orth <= min(WARN_U, 10.0 * n))}


def gather_to_rank0(A_loc, Q_loc, n, ld, layout_of, dist):
    """Collective (one process per GPU): H (block-cyclic columns) and Q (row slabs) of all ranks on rank 0's GPU.
    layout_of(r) -> starneig_b200.dist.Layout of rank r. Returns (Ht, Qt) on rank 0, (None, None) elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = A_loc.device
    if rank != 0:
        dist.send(A_loc, dst=0)
        dist.send(Q_loc, dst=0)
        return None, None
    H = torch.empty((n, ld), dtype=torch.float64, device=dev)
    Qf = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    L0 = layout_of(0)
    H[torch.from_numpy(L0.global_cols()).to(dev)] = A_loc
    Qf[:, L0.q_row0:L0.q_row0 + L0.q_rows] = Q_loc[:, :L0.q_rows]
    for r in range(1, world):
        Lr = layout_of(r)
        buf = torch.empty((Lr.local_cols, ld), dtype=torch.float64, device=dev)
        dist.recv(buf, src=r)
        H[torch.from_numpy(Lr.global_cols()).to(dev)] = buf
        del buf
        ldq_r = (max(Lr.q_rows, 1) + 15) // 16 * 16
        buf = torch.empty((n, ldq_r), dtype=torch.float64, device=dev)
        dist.recv(buf, src=r)
        Qf[:, Lr.q_row0:Lr.q_row0 + Lr.q_rows] = buf[:, :Lr.q_rows]
        del buf
    return H, Qf
