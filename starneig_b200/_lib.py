"""ctypes binding of the C-ABI library ``lib/libstarneig.so`` (built by ``csrc/Makefile``).

There is no Python or CPU fallback: if the CUDA library is missing, importing the binding raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstarneig.so")

c_double_p = ctypes.POINTER(ctypes.c_double)


class HessenbergConf(ctypes.Structure):
    """``struct starneig_hessenberg_conf`` (reference src/include/starneig/expert.h:77-92)."""
    _fields_ = [("tile_size", ctypes.c_int), ("panel_width", ctypes.c_int)]


class Stats(ctypes.Structure):
    """``struct starneig_b200_stats`` (include/starneig_b200.h)."""
    _fields_ = [
        ("n", ctypes.c_int), ("begin", ctypes.c_int), ("end", ctypes.c_int),
        ("panel_width", ctypes.c_int), ("panels", ctypes.c_int),
        ("wall_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double),
        ("device_ms", ctypes.c_double), ("panel_ms", ctypes.c_double), ("trail_ms", ctypes.c_double),
        ("other_ms", ctypes.c_double), ("gemv_ms", ctypes.c_double),
        ("gemv_launches", ctypes.c_longlong), ("gemv_bytes", ctypes.c_double),
        ("gemv_timed_launches", ctypes.c_longlong), ("gemv_timed_bytes", ctypes.c_double),
        ("finish_update_ms", ctypes.c_double), ("reflector_ms", ctypes.c_double),
        ("kernel_launches", ctypes.c_longlong), ("gemm_flops", ctypes.c_double),
        ("h2d_bytes", ctypes.c_longlong), ("d2h_bytes", ctypes.c_longlong),
        ("ranks", ctypes.c_int), ("fused_panels", ctypes.c_int), ("fused_kernel_ms", ctypes.c_double),
        ("fused_phase_ms", ctypes.c_double * 4),
        ("overlap", ctypes.c_int), ("side_tail_ms", ctypes.c_double),
        ("gemm_tma_launches", ctypes.c_longlong), ("gemm_cpasync_launches", ctypes.c_longlong),
        ("staging_overlapped", ctypes.c_int), ("panel_width_used", ctypes.c_int),
        ("q_backward", ctypes.c_int), ("q_backward_ms", ctypes.c_double),
        ("fused_slab_panels", ctypes.c_int * 2),
    ]

    def as_dict(self):
        return {name: (list(getattr(self, name)) if name in ("fused_phase_ms", "fused_slab_panels") else getattr(self, name)) for name, _ in self._fields_}


SCHUR_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                            ctypes.c_void_p, ctypes.c_void_p)
PREDICATE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p)
SELECT_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, PREDICATE_FN, ctypes.c_void_p,
                             ctypes.c_void_p, ctypes.POINTER(ctypes.c_int))
REORDER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)


class Chain(ctypes.Structure):
    """``struct starneig_b200_chain`` (include/starneig_b200.h): the next stages of a Reduce-shaped chain."""
    _fields_ = [("schur", SCHUR_FN), ("schur_device", SCHUR_FN), ("select", SELECT_FN), ("reorder_schur", REORDER_FN)]


def load(path=None):
    """Binds the C ABI. `path` is only given by tests/test_cusim.py, which binds the same entry points of the
    kernel-logic emulator build of the library (tests/cusim); the product always loads LIB_PATH."""
    if path is None:
        # STARNEIG_B200_LIB: another build of the same sources (A/B experiments on hardware, csrc/Makefile `exp`)
        path = os.environ.get("STARNEIG_B200_LIB") or LIB_PATH
        if not os.path.exists(path):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C starneig_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU fallback for the Hessenberg path.")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    i, d, vp = ctypes.c_int, ctypes.c_double, ctypes.c_void_p

    lib.starneig_node_init.argtypes = [i, i, ctypes.c_uint]
    lib.starneig_node_init.restype = None
    lib.starneig_node_initialized.restype = i
    lib.starneig_node_finalize.restype = None
    lib.starneig_node_get_cores.restype = i
    lib.starneig_node_get_gpus.restype = i
    lib.starneig_node_set_cores.argtypes = [i]
    lib.starneig_node_set_gpus.argtypes = [i]
    lib.starneig_node_enable_pinning.restype = None
    lib.starneig_node_disable_pinning.restype = None
    lib.starneig_hessenberg_init_conf.argtypes = [ctypes.POINTER(HessenbergConf)]
    lib.starneig_hessenberg_init_conf.restype = None
    lib.starneig_SEP_SM_Hessenberg.argtypes = [i, vp, i, vp, i]
    lib.starneig_SEP_SM_Hessenberg.restype = i
    lib.starneig_SEP_SM_Hessenberg_expert.argtypes = [ctypes.POINTER(HessenbergConf), i, i, i, vp, i, vp, i]
    lib.starneig_SEP_SM_Hessenberg_expert.restype = i
    lib.starneig_b200_hessenberg_device.argtypes = [i, i, i, i, vp, i, vp, i]
    lib.starneig_b200_hessenberg_device.restype = i
    lib.starneig_b200_get_stats.argtypes = [ctypes.POINTER(Stats)]
    lib.starneig_b200_get_stats.restype = None
    lib.starneig_b200_set_profile_level.argtypes = [i]
    lib.starneig_b200_set_profile_level.restype = None
    lib.starneig_b200_dgemm.argtypes = [ctypes.c_char, ctypes.c_char, i, i, i, d, vp, i, vp, i, d, vp, i]
    lib.starneig_b200_dgemm.restype = i
    lib.starneig_b200_gemv.argtypes = [i, i, vp, i, vp, vp, i, ctypes.POINTER(ctypes.c_float)]
    lib.starneig_b200_gemv.restype = i
    lib.starneig_b200_panel.argtypes = [i, i, i, i, vp, i, vp, vp, vp, i, vp]
    lib.starneig_b200_panel.restype = i
    ip = ctypes.POINTER(ctypes.c_int)
    lib.starneig_b200_dist_layout.argtypes = [i, i, i, ip, ip, ip, ip]
    lib.starneig_b200_dist_layout.restype = i
    lib.starneig_b200_dist_global_col.argtypes = [i, i, i, i]
    lib.starneig_b200_dist_global_col.restype = i
    lib.starneig_b200_dist_init.argtypes = [i, i, i, i, vp]
    lib.starneig_b200_dist_init.restype = i
    lib.starneig_b200_dist_connect.argtypes = [vp]
    lib.starneig_b200_dist_connect.restype = i
    lib.starneig_b200_dist_hessenberg_device.argtypes = [i, i, i, i, vp, i, vp, i]
    lib.starneig_b200_dist_hessenberg_device.restype = i
    lib.starneig_b200_dist_hessenberg_host.argtypes = [i, i, i, i, vp, i, vp, i]
    lib.starneig_b200_dist_hessenberg_host.restype = i
    lib.starneig_b200_dist_finalize.restype = None
    lib.starneig_b200_SEP_SM_Hessenberg_stage.argtypes = [i, vp, i, vp, i, ctypes.POINTER(vp), ip, ctypes.POINTER(vp), ip]
    lib.starneig_b200_SEP_SM_Hessenberg_stage.restype = i
    lib.starneig_b200_stage_fetch.argtypes = [i, vp, i, vp, i]
    lib.starneig_b200_stage_fetch.restype = i
    lib.starneig_b200_SEP_SM_Reduce.argtypes = [i, vp, i, vp, i, vp, vp, PREDICATE_FN, vp, vp, ip, ctypes.POINTER(Chain)]
    lib.starneig_b200_SEP_SM_Reduce.restype = i
    lib.starneig_b200_plan_check.argtypes = [i, i, i, ctypes.POINTER(ctypes.c_longlong)]
    lib.starneig_b200_plan_check.restype = i
    return lib
