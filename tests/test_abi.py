"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the headers declare
(and nothing else), and reproduces the reference's argument/state error behaviour. No compute calls."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = set()
    for path in [os.path.join(INCLUDE, "starneig_b200.h")] + [
            os.path.join(INCLUDE, "starneig", f) for f in sorted(os.listdir(os.path.join(INCLUDE, "starneig")))]:
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(starneig_[A-Za-z0-9_]+)\s*\(", text))
    return {x for x in names if not x.endswith("_t")}       # (a type in front of a function-pointer member is not a function)


def exported_symbols():
    from starneig_b200._lib import LIB_PATH
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], check=True, capture_output=True, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_exports_match_headers():
    decl, exp = declared_symbols(), exported_symbols()
    assert decl, "no declarations found"
    missing = decl - exp
    assert not missing, f"declared in include/ but not exported: {sorted(missing)}"
    # everything else is hidden, as in the reference (src/CMakeLists.txt:154-155)
    extra = {s for s in exp if not s.startswith("starneig_")}
    assert not extra, f"unexpected exported symbols: {sorted(extra)}"


def test_headers_compile_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <starneig/starneig.h>\n#include <starneig_b200.h>\n'
                   'int main(void){struct starneig_hessenberg_conf c; (void)c; return STARNEIG_SUCCESS;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INCLUDE, "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_c_program_links_against_library(tmp_path):
    # the reference's own README sample shape: node_init / Hessenberg / node_finalize (README.md:192-199)
    from starneig_b200._lib import LIB_PATH
    src = tmp_path / "t.c"
    src.write_text('#include <starneig/starneig.h>\n#include <stdio.h>\n'
                   'int main(void){ double A[4]={1,2,3,4}, Q[4]={1,0,0,1};\n'
                   ' int r0 = starneig_SEP_SM_Hessenberg(2, A, 2, Q, 2);\n'
                   ' starneig_node_init(1, 0, STARNEIG_NO_MESSAGES);\n'
                   ' int r1 = starneig_SEP_SM_Hessenberg(0, A, 2, Q, 2);\n'
                   ' int r2 = starneig_SEP_SM_Hessenberg_expert(0, 2, 0, 3, A, 2, Q, 2);\n'
                   ' starneig_node_finalize(); printf("%d %d %d\\n", r0, r1, r2); return 0; }\n')
    exe = tmp_path / "t"
    libdir = os.path.dirname(LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", INCLUDE, str(src), "-o", str(exe), "-L", libdir, "-l:libstarneig.so",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["2", "-1", "-4"]


def test_error_codes_and_state_machine(sn):
    n = 8
    A = np.zeros((n, n), order="F"); Q = np.asfortranarray(np.eye(n))
    # before init: argument errors win over NOT_INITIALIZED (interface.c:175-182)
    assert not sn.starneig_node_initialized()
    assert sn.starneig_SEP_SM_Hessenberg(0, A, n, Q, n) == -1
    assert sn.starneig_SEP_SM_Hessenberg(n, A, n, Q, n) == sn.STARNEIG_NOT_INITIALIZED
    assert sn.starneig_SEP_SM_Hessenberg_expert(None, n, 0, n, A, n, Q, n) == sn.STARNEIG_NOT_INITIALIZED
    sn.starneig_node_init(1, 0, sn.STARNEIG_NO_MESSAGES)
    try:
        assert sn.starneig_node_initialized()
        assert sn.starneig_node_get_cores() == 1
        assert sn.starneig_node_get_gpus() == 0
        # simple interface, interface.c:175-179
        assert sn.starneig_SEP_SM_Hessenberg(0, A, n, Q, n) == -1
        assert sn.starneig_SEP_SM_Hessenberg(n, None, n, Q, n) == -2
        assert sn.starneig_SEP_SM_Hessenberg(n, A, n - 1, Q, n) == -3
        assert sn.starneig_SEP_SM_Hessenberg(n, A, n, None, n) == -4
        assert sn.starneig_SEP_SM_Hessenberg(n, A, n, Q, n - 1) == -5
        # expert interface, interface.c:144-150
        conf = sn.starneig_hessenberg_init_conf()
        assert (conf.tile_size, conf.panel_width) == (-1, -1)
        E = sn.starneig_SEP_SM_Hessenberg_expert
        assert E(conf, 0, 0, 0, A, n, Q, n) == -2
        assert E(conf, n, -1, n, A, n, Q, n) == -3
        assert E(conf, n, 0, n + 1, A, n, Q, n) == -4
        assert E(conf, n, 0, n, None, n, Q, n) == -5
        assert E(conf, n, 0, n, A, n - 1, Q, n) == -6
        assert E(conf, n, 0, n, A, n, None, n) == -7
        assert E(conf, n, 0, n, A, n, Q, n - 1) == -8
        # invalid configuration, interface.c:67-84
        conf.panel_width = 4
        assert E(conf, n, 0, n, A, n, Q, n) == sn.STARNEIG_INVALID_CONFIGURATION
        conf.panel_width = -1; conf.tile_size = 7
        assert E(conf, n, 0, n, A, n, Q, n) == sn.STARNEIG_INVALID_CONFIGURATION
        # no GPU selected: the CUDA-only path refuses loudly, it never falls back to a CPU implementation
        conf.tile_size = -1
        assert E(conf, n, 0, n, A, n, Q, n) == sn.STARNEIG_GENERIC_ERROR
        assert np.all(A == 0.0)
    finally:
        sn.starneig_node_finalize()
    assert not sn.starneig_node_initialized()


def test_double_init_is_fatal(tmp_path):
    # reference: "The node is already initialized." -> exit(EXIT_FAILURE) (node.c:442-443, common.c:143-152)
    code = ("import sys; sys.path.insert(0, %r); import starneig_b200 as s; "
            "s.starneig_node_init(1,0,0x30); s.starneig_node_init(1,0,0x30)") % ROOT
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "[starneig][fatal error] The node is already initialized." in r.stderr


def test_product_never_touches_the_oracle():
    # a product path that routes through the oracle would void every parity claim
    pkg = os.path.join(ROOT, "starneig_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f"{f} mentions the oracle"
    from starneig_b200._lib import LIB_PATH
    needed = subprocess.run(["readelf", "-d", LIB_PATH], capture_output=True, text=True).stdout
    assert "openblas" not in needed.lower() and "liboracle" not in needed.lower()
    # ... nor the kernel-logic emulator of the test suite (tests/cusim): the shipped library is the nvcc build, it
    # links the real CUDA runtime and contains none of the emulator's symbols
    assert "libcudart" in needed.lower() and "starneig_sim" not in needed.lower()
    syms = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    assert "cusim" not in syms.lower()
    assert "__cudaRegisterFatBinary" in subprocess.run(["nm", "-D", LIB_PATH], capture_output=True, text=True).stdout
    # the C test driver is a client of the library and of a CPU BLAS (for its residual checks), never of the oracle
    driver = os.path.join(ROOT, "driver", "bin", "starneig-test")
    if os.path.exists(driver):
        needed = subprocess.run(["readelf", "-d", driver], capture_output=True, text=True).stdout
        assert "liboracle" not in needed.lower() and "libstarneig_ref" not in needed.lower()


def test_workspace_plan_bounds_up_to_the_largest_supported_order(sn):
    """Host-only arithmetic of both panel paths (starneig_b200_plan_check): for every size up to STARNEIG_B200_MAX_N the
    GEMV partial sums of any column fit the buffer that is allocated for them, a panel width that fits the panel
    kernels' shared memory is found (narrower than the reference default beyond n ~ 70000), and larger n is rejected."""
    import ctypes
    lib = sn.lib()
    out = (ctypes.c_longlong * 4)()
    MAX_N = 131056
    sizes = [1, 2, 9, 300, 4000, 20000, 46000, 47500, 50000, 60000, 75000, 76001, 80000, 100000, 120000, MAX_N]
    for n in sizes:
        for ranks in (1, 2, 8):
            for pw in (-1, 64, 512, 1024):
                assert lib.starneig_b200_plan_check(n, pw, ranks, out) == 0, (n, ranks, pw)
                used, smem, cap, worst = out[0], out[1], out[2], out[3]
                assert 8 <= used <= 1024 and (pw < 0 or used <= max(pw, 8))
                assert smem <= 200 * 1024
                assert worst <= cap, (n, ranks, pw, worst, cap)
    # the automatic width (192: two 96-column tiles of the skinny DMMA products, measured on B200) at any size and GPU count
    for n, ranks in ((20000, 1), (50000, 1), (20000, 2), (20000, 4), (20000, 8), (50000, 8)):
        assert lib.starneig_b200_plan_check(n, -1, ranks, out) == 0 and out[0] == 192 and out[1] > 0
    # the reference's own default (312 / 368) is still held by the persistent kernel's layout when a caller asks for it
    assert lib.starneig_b200_plan_check(20000, 312, 1, out) == 0 and out[0] == 312 and out[1] > 0
    assert lib.starneig_b200_plan_check(50000, 368, 1, out) == 0 and out[0] == 368 and out[1] > 0
    assert lib.starneig_b200_plan_check(100000, -1, 1, out) == 0 and out[0] < sn.default_panel_width(100000)
    assert lib.starneig_b200_plan_check(MAX_N + 1, -1, 1, out) == 4          # STARNEIG_INVALID_ARGUMENTS
    assert lib.starneig_b200_plan_check(0, -1, 1, out) == 4


def test_dmma_fragment_loads_are_bank_conflict_free():
    """the k permutation of the TMA-fed DMMA kernels (dgemm_tma.cuh::tma_kperm) against the 128-byte swizzle pattern: brute force
    over every warp position, step and half-warp for both operand layouts (tools/swizzle_check.py)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("swizzle_check", os.path.join(ROOT, "tools", "swizzle_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    natural = [[4 * s + t for t in range(4)] for s in range(4)]
    assert mod.worst(mod.addr_mn, natural) == 2 and mod.worst(mod.addr_k, natural) == 2
    assert mod.worst(mod.addr_mn, mod.KSETS) == 1 and mod.worst(mod.addr_k, mod.KSETS) == 1
    # the header's table is the one checked here
    src = open(os.path.join(ROOT, "starneig_b200", "csrc", "dgemm_tma.cuh")).read()
    for s, ks in enumerate(mod.KSETS):
        assert "t == 0 ? %d : t == 1 ? %d : t == 2 ? %d : %d" % tuple(ks) in src, (s, ks)
