"""The test-driver-compatible harness (driver/starneig_test.c): option handling, the
reference driver's generators and raw matrix format on the CPU (through its `lapack` comparison solver), and --
on a GPU -- the product library behind `--solver starneig` / `starneig-simple` against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "driver", "bin", "starneig-test")
U = 2.0 ** -52


@pytest.fixture(scope="module")
def driver():
    subprocess.run(["make", "-C", os.path.join(ROOT, "driver")], check=True, capture_output=True)
    assert os.path.exists(DRIVER)
    return DRIVER


def run(driver, *args, cwd=None, ok=(0,)):
    p = subprocess.run([driver, *map(str, args)], capture_output=True, text=True, cwd=cwd, timeout=600)
    assert p.returncode in ok, p.stdout + p.stderr
    return p.stdout


def read_raw(path):
    """reference raw format (test/common/io.c:236-360): text header line, then column-major doubles"""
    with open(path, "rb") as f:
        header = f.readline().decode()
        m = re.fullmatch(r"STARNEIG RAW REAL DOUBLE M (\d+) N (\d+)\n", header)
        assert m, header
        rows, cols = int(m.group(1)), int(m.group(2))
        data = np.frombuffer(f.read(), dtype=np.float64)
    assert data.size == rows * cols
    return data.reshape((cols, rows)).T.copy(order="F")


def test_usage_and_bad_arguments(driver):
    assert "Usage" in run(driver, ok=(2,))
    run(driver, "--experiment", "hessenberg", "--solver", "nope", ok=(2,))
    run(driver, "--experiment", "hessenberg", "--n", 0, "--solver", "lapack", ok=(2,))
    run(driver, "--experiment", "hessenberg", "--n", 10, "--begin", 5, "--end", 3, "--solver", "lapack", ok=(2,))


def test_generators_checks_and_raw_format_on_cpu(driver, ora, tmp_path):
    n, seed = 123, 2019
    out = run(driver, "--experiment", "hessenberg", "--n", n, "--seed", seed, "--solver", "lapack",
              "--hooks", "hessenberg", "residual", "store-raw", "--store-raw-output", tmp_path / "h_%s.dat")
    assert "NO FAILED HESSENBERG FORM TESTS" in out and "EXPERIMENT TIME" in out
    assert re.search(r"\|Q ~A Q\^T - A\| / \|A\| = \d+ u", out) and re.search(r"\|Q Q\^T - I\| / \|I\| = \d+ u", out)
    CA, H, Q = (read_raw(tmp_path / f"h_{t}.dat") for t in ("CA", "A", "Q"))
    # the driver's LCG 'fullpos' initialiser == the reference driver's (oracle restatement, pinned by golden fixtures)
    A0, _, ld = ora.fullpos(n, seed)
    assert np.array_equal(CA, A0[:n])
    assert np.count_nonzero(np.tril(H, -2)) == 0
    # dgehrd/dormhr result == the oracle's LAPACK solver on the same input
    A2, Q2 = A0.copy(order="F"), np.zeros_like(A0); Q2[:n] = np.eye(n)
    assert ora.hessenberg_lapack(n, A2, ld, Q2, ld) == 0
    assert np.abs(H - A2[:n]).max() <= 100 * n * U * np.abs(A2[:n]).max() and np.abs(Q - Q2[:n]).max() <= 100 * n * U

    # read-raw round trip: the stored input reproduces the stored output bit for bit
    out = run(driver, "--experiment", "hessenberg", "--init", "read-raw", "--input", tmp_path / "h_C%s.dat", "--solver", "lapack",
              "--hooks", "hessenberg", "store-raw", "--store-raw-output", tmp_path / "again_%s.dat")
    assert f"READING A {n} X {n} MATRIX" in out
    assert np.array_equal(read_raw(tmp_path / "again_A.dat"), H)


def test_full_and_partial_initialisers(driver, ora, tmp_path):
    n, seed = 88, 7
    run(driver, "--experiment", "hessenberg", "--init", "full", "--n", n, "--seed", seed, "--solver", "lapack",
        "--hooks", "store-raw", "--store-raw-output", tmp_path / "f_%s.dat")
    assert np.array_equal(read_raw(tmp_path / "f_CA.dat"), ora.full(n, seed)[0][:n])
    out = run(driver, "--experiment", "partial-hessenberg", "--n", n, "--seed", seed, "--solver", "lapack",
              "--hooks", "hessenberg", "residual", "store-raw", "--store-raw-output", tmp_path / "p_%s.dat")
    assert f"--begin {n // 4} --end {3 * n // 4}" in out and "NO FAILED HESSENBERG FORM TESTS" in out
    assert np.array_equal(read_raw(tmp_path / "p_CA.dat"), ora.partial(n, n // 4, 3 * n // 4, seed)[0][:n])


def test_without_gpu_the_product_solver_fails_loudly(driver):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    p = subprocess.run([driver, "--experiment", "hessenberg", "--n", "50", "--pinning", "off"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("solver,extra", [("starneig", ["--panel-width", 96]), ("starneig-simple", [])])
def test_product_solver_against_oracle(driver, ora, tmp_path, solver, extra):
    n, seed = 700, 2019
    out = run(driver, "--experiment", "hessenberg", "--n", n, "--seed", seed, "--solver", solver, "--gpus", 1, *extra,
              "--hooks", "hessenberg", "residual", "store-raw", "--store-raw-output", tmp_path / "g_%s.dat",
              "--repeat", 2, "--warmup", 1)
    assert "NO FAILED HESSENBERG FORM TESTS" in out and "FAILS" not in out and "WARNINGS" not in out
    H, Q = read_raw(tmp_path / "g_A.dat"), read_raw(tmp_path / "g_Q.dat")
    A0, Q0, ld = ora.fullpos(n, seed)
    A2, Q2 = A0.copy(order="F"), Q0.copy(order="F")
    assert ora.hessenberg_port(n, A2, ld, Q2, ld, 0, n, 96 if extra else -1) == 0
    assert np.abs(H - A2[:n]).max() <= 200 * n * U * np.abs(A2[:n]).max()
    assert np.abs(Q - Q2[:n]).max() <= 200 * n * U
    assert np.array_equal(H == 0.0, A2[:n] == 0.0)


@pytest.mark.gpu
def test_product_solver_partial_experiment(driver):
    out = run(driver, "--experiment", "partial-hessenberg", "--n", 554, "--seed", 11, "--solver", "starneig", "--gpus", 1,
              "--panel-width", 45)
    assert "NO FAILED HESSENBERG FORM TESTS" in out and "FAILS" not in out
