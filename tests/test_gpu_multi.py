"""GPU parity tests of the multi-GPU engine: 1-D block-cyclic column shards of A, row slabs of Q,
per-column exchange of the GEMV sums over peer memory.

Two drivers of the same engine are covered:
  * one process, `gpus` host threads (what starneig_node_init(cores, gpus, ...) + starneig_SEP_SM_Hessenberg do);
  * one process per GPU (torchrun + CUDA IPC), through tests/dist_worker.py.
On a box with fewer GPUs than ranks the ranks share devices (STARNEIG_B200_VIRTUAL_RANKS, a development aid): the
code path -- flags, peer stores, barriers -- is the same, only the transport is local HBM instead of NVLink.
Results must equal the single-GPU result to rounding (the engines sum the same partials in a different order) and
satisfy the oracle comparison with the same tolerances as tests/test_gpu_hessenberg.py.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, same_zero_pattern, structured_input

pytestmark = pytest.mark.gpu
U = 2.0 ** -52


@pytest.fixture()
def team(sn, monkeypatch):
    """node factory: team(P) initialises the node with P GPUs (virtual ranks if the box has fewer)"""
    monkeypatch.setenv("STARNEIG_B200_VIRTUAL_RANKS", "8")
    state = {"init": False}

    def make(P):
        if state["init"]:
            sn.starneig_node_finalize()
        sn.starneig_node_init(sn.STARNEIG_USE_ALL, P, sn.STARNEIG_NO_MESSAGES)
        state["init"] = True
        assert sn.starneig_node_get_gpus() == P
        return sn
    yield make
    if state["init"]:
        sn.starneig_node_finalize()


def _run(sn, n, A, ld, Q, begin=0, end=None, pw=-1):
    conf = sn.starneig_hessenberg_init_conf()
    conf.panel_width = pw
    return sn.starneig_SEP_SM_Hessenberg_expert(conf, n, begin, n if end is None else end, A, ld, Q, ld)


def _check(ora, n, A, Q, A0, Q0, ld, begin=0, end=None, pw=-1):
    end = n if end is None else end
    assert np.isfinite(A[:n]).all() and np.isfinite(Q[:n]).all()
    assert ora.hessenberg_form_violations(n, A, ld, begin, end, check_outside=True) == 0
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, begin, end, pw) == 0
    assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * max(1.0, np.abs(A2[:n]).max())
    assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U
    assert same_zero_pattern(A, A2, n)
    res = ora.residual_u(n, Q, ld, A, ld, A0, ld)
    orth = ora.orthogonality_u(n, Q, ld)
    assert res <= max(10.0 * n, 20.0) and res <= 500 and orth <= max(10.0 * n, 20.0) and orth <= 500, (res, orth)


@pytest.mark.parametrize("P", [2, 3, 4, 8])
@pytest.mark.parametrize("n,pw", [(2, 8), (9, 8), (47, 16), (130, 35), (333, 45), (700, 170), (1100, -1)])
def test_threads_against_oracle(team, ora, P, n, pw):
    sn = team(P)
    A0, Q0, ld = ora.fullpos(n, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(sn, n, A, ld, Q, pw=pw) == 0
    assert sn.get_stats()["ranks"] == P
    _check(ora, n, A, Q, A0, Q0, ld, pw=pw)


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("n", [88, 333, 554])
def test_threads_partial_reduction(team, ora, P, n):
    sn = team(P)
    begin, end = n // 4, 3 * n // 4
    A0, Q0, ld = ora.partial(n, begin, end, 2019)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(sn, n, A, ld, Q, begin, end, pw=16) == 0
    _check(ora, n, A, Q, A0, Q0, ld, begin, end, pw=16)


def test_threads_col_block_sizes(team, ora, monkeypatch):
    """column-block widths that do and do not divide the panel width; non-identity Q; ld > n"""
    n, pw = 500, 48
    A0, Q0, ld = ora.full(n, 7)
    rng = np.random.default_rng(3)
    Qr, _ = np.linalg.qr(rng.standard_normal((n, n)))
    Q0[:n, :n] = Qr
    for cb in (8, 24, 64, 200):
        monkeypatch.setenv("STARNEIG_B200_COL_BLOCK", str(cb))
        sn = team(3)
        A, Q = A0.copy(order="F"), Q0.copy(order="F")
        assert _run(sn, n, A, ld, Q, pw=pw) == 0
        A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
        assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, pw) == 0
        assert np.abs(A[:n] - A2[:n]).max() <= 200 * n * U * max(1.0, np.abs(A2[:n]).max())
        assert np.abs(Q[:n] - Q2[:n]).max() <= 200 * n * U


def test_threads_match_single_gpu_and_repeat(team, ora):
    """P ranks vs 1 rank on the same input: same exact-zero pattern, entries equal to rounding; two consecutive
    multi-rank calls (arena reuse, epochs continue) are bitwise identical"""
    n = 900
    A0, Q0, ld = ora.fullpos(n, 11)
    sn = team(1)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(sn, n, A1, ld, Q1) == 0
    sn = team(4)
    outs = []
    for _ in range(2):
        A, Q = A0.copy(order="F"), Q0.copy(order="F")
        assert _run(sn, n, A, ld, Q) == 0
        outs.append((A, Q))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][0][:n] - A1[:n]).max() <= 50 * n * U * np.abs(A1[:n]).max()
    assert np.abs(outs[0][1][:n] - Q1[:n]).max() <= 50 * n * U
    assert same_zero_pattern(outs[0][0], A1, n)


# columns in which DLARFG meets x = 0 (tau = 0) while the GEMV sums of all ranks are exchanged (tests/conftest.py)
@pytest.mark.parametrize("name", ["upper_triangular", "zero_columns", "block_triangular"])
@pytest.mark.parametrize("P", [2, 4])
def test_threads_structured_inputs(team, ora, P, name):
    n, pw = 200, 24
    sn = team(P)
    A0, Q0, ld, _ = structured_input(ora, name, n)
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(sn, n, A, ld, Q, pw=pw) == 0
    _check(ora, n, A, Q, A0, Q0, ld, pw=pw)


@pytest.mark.parametrize("world", [2, 4])
def test_processes_torchrun(world, tmp_path):
    """one process per GPU: torchrun + NCCL for the handle exchange, CUDA IPC peer memory for the data path"""
    import torch
    env = dict(os.environ)
    env["STARNEIG_B200_VIRTUAL_RANKS"] = "8"
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    port = 29500 + (os.getpid() % 500) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), "--size", "600", "--panel", "40",
           "--devices", str(torch.cuda.device_count())]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_WORKER_OK" in out.stdout


def test_real_peers_at_size(team, ora):
    """every visible GPU as one rank (real NVLink peers where the box has several), n = 4000, default (per-GPU-count) panel
    width: the reference driver's acceptance checks on the whole result and agreement with the one-GPU result"""
    import torch
    from tools import invariants
    P = max(1, min(8, torch.cuda.device_count()))
    n = 4000
    A0, Q0, ld = ora.fullpos(n, 2019)
    A1, Q1 = A0.copy(order="F"), Q0.copy(order="F")
    assert _run(team(1), n, A1, ld, Q1) == 0
    A, Q = A0.copy(order="F"), Q0.copy(order="F")
    sn = team(P)
    assert _run(sn, n, A, ld, Q) == 0
    st = sn.get_stats()
    assert st["ranks"] == P
    t = lambda M: torch.from_numpy(np.ascontiguousarray(M.T)).cuda()
    inv = invariants.evaluate(t(A0), t(A), t(Q), n)
    assert inv["ok"], inv
    assert np.abs(A[:n] - A1[:n]).max() <= 200 * n * U * np.abs(A1[:n]).max() and np.abs(Q[:n] - Q1[:n]).max() <= 200 * n * U
    print(f"real peers: {P} GPUs, n = {n}, panel width {st['panel_width_used']}: {inv}")
