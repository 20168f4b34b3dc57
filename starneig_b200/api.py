"""Host-side mirror of the reference's public interface for the Hessenberg path.

Same names, argument meaning and error behaviour as ``<starneig/node.h>`` and ``<starneig/sep_sm.h>``
(reference src/include/starneig/node.h:178-241, sep_sm.h:89-92,380-384); every call goes straight
through the C ABI of ``libstarneig.so``. NumPy arrays stand for the caller's column-major host
buffers; torch CUDA tensors for device-resident matrices.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import HessenbergConf, Stats

# error codes, reference src/include/starneig/error.h:66-127
STARNEIG_SUCCESS = 0
STARNEIG_GENERIC_ERROR = 1
STARNEIG_NOT_INITIALIZED = 2
STARNEIG_INVALID_CONFIGURATION = 3
STARNEIG_INVALID_ARGUMENTS = 4

# init flags, reference src/include/starneig/node.h:78-158
STARNEIG_DEFAULT = 0x0
STARNEIG_HINT_SM = 0x0
STARNEIG_HINT_DM = 0x1
STARNEIG_FXT_DISABLE = 0x2
STARNEIG_AWAKE_WORKERS = 0x4
STARNEIG_AWAKE_MPI_WORKER = 0x8
STARNEIG_NO_VERBOSE = 0x10
STARNEIG_NO_MESSAGES = 0x30
STARNEIG_USE_ALL = -1

STARNEIG_HESSENBERG_DEFAULT_TILE_SIZE = -1
STARNEIG_HESSENBERG_DEFAULT_PANEL_WIDTH = -1

_handle = None


def lib():
    global _handle
    if _handle is None:
        _handle = _lib.load()
    return _handle


def starneig_node_init(cores=STARNEIG_USE_ALL, gpus=STARNEIG_USE_ALL, flags=STARNEIG_DEFAULT):
    lib().starneig_node_init(cores, gpus, flags)


def starneig_node_initialized():
    return bool(lib().starneig_node_initialized())


def starneig_node_finalize():
    lib().starneig_node_finalize()


def starneig_node_get_cores():
    return lib().starneig_node_get_cores()


def starneig_node_get_gpus():
    return lib().starneig_node_get_gpus()


def starneig_node_set_cores(cores):
    lib().starneig_node_set_cores(cores)


def starneig_node_set_gpus(gpus):
    lib().starneig_node_set_gpus(gpus)


def starneig_node_enable_pinning():
    lib().starneig_node_enable_pinning()


def starneig_node_disable_pinning():
    lib().starneig_node_disable_pinning()


def starneig_hessenberg_init_conf():
    conf = HessenbergConf()
    lib().starneig_hessenberg_init_conf(ctypes.byref(conf))
    return conf


def _host_ptr(a):
    if a is None:
        return None
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.f_contiguous:
        raise TypeError("expected a column-major (order='F') float64 numpy array")
    return a.ctypes.data


def starneig_SEP_SM_Hessenberg(n, A, ldA, Q, ldQ):
    """A <- H, Q <- Q*U in place; returns the ``starneig_error_t`` of the C call."""
    return lib().starneig_SEP_SM_Hessenberg(n, _host_ptr(A), ldA, _host_ptr(Q), ldQ)


def starneig_SEP_SM_Hessenberg_expert(conf, n, begin, end, A, ldA, Q, ldQ):
    cref = ctypes.byref(conf) if conf is not None else None
    return lib().starneig_SEP_SM_Hessenberg_expert(cref, n, begin, end, _host_ptr(A), ldA, _host_ptr(Q), ldQ)


def hessenberg_device(n, A, ldA, Q, ldQ, begin=0, end=None, panel_width=-1):
    """Device-resident variant: A, Q are column-major float64 CUDA buffers (torch tensors)."""
    end = n if end is None else end
    return lib().starneig_b200_hessenberg_device(n, begin, end, panel_width, A.data_ptr(), ldA, Q.data_ptr(), ldQ)


def hessenberg_stage(n, A, ldA, Q, ldQ):
    """``starneig_b200_SEP_SM_Hessenberg_stage``: the Hessenberg stage of a chain whose next stage runs on the GPU. Host A, Q are
    inputs only; returns (ret, dH, lddH, dQ, lddQ) with dH, dQ the device addresses (ints) of H and Q."""
    dH, dQ = ctypes.c_void_p(), ctypes.c_void_p()
    lddH, lddQ = ctypes.c_int(), ctypes.c_int()
    ret = lib().starneig_b200_SEP_SM_Hessenberg_stage(n, _host_ptr(A), ldA, _host_ptr(Q), ldQ, ctypes.byref(dH), ctypes.byref(lddH),
                                                      ctypes.byref(dQ), ctypes.byref(lddQ))
    return ret, dH.value, lddH.value, dQ.value, lddQ.value


def stage_fetch(n, A, ldA, Q, ldQ):
    """``starneig_b200_stage_fetch``: H and Q of the last stage call -> host arrays (either may be None)."""
    return lib().starneig_b200_stage_fetch(n, _host_ptr(A), ldA, _host_ptr(Q), ldQ)


def starneig_b200_SEP_SM_Reduce(n, A, ldA, Q, ldQ, real, imag, chain, predicate=None, arg=None, selected=None):
    """``starneig_b200_SEP_SM_Reduce`` (shape of the reference's starneig_SEP_SM_Reduce, src/common/combined.c:45-98).
    `chain`: a ``_lib.Chain`` of ctypes callbacks for the stages this library does not own. Returns (ret, num_selected)."""
    num = ctypes.c_int(0)
    pred = predicate if predicate is not None else _lib.PREDICATE_FN()
    sel = selected.ctypes.data if selected is not None else None
    ret = lib().starneig_b200_SEP_SM_Reduce(n, _host_ptr(A), ldA, _host_ptr(Q), ldQ, real.ctypes.data, imag.ctypes.data, pred, arg, sel,
                                            ctypes.byref(num), ctypes.byref(chain))
    return ret, num.value


def get_stats():
    st = Stats()
    lib().starneig_b200_get_stats(ctypes.byref(st))
    return st.as_dict()


def set_profile_level(level):
    lib().starneig_b200_set_profile_level(level)


def default_panel_width(n):
    """the REFERENCE's automatic panel width (src/hessenberg/interface.c:74-78), e.g. to run both with one blocking; the
    library's own automatic width is 192 (engine.cuh::default_panel_width, measured on B200)"""
    import math
    return max(64, int(math.ceil((0.001875596476 * n + 273.5908216) / 8.0)) * 8)
