"""starneig_b200 -- B200-native (sm_100a) blocked Hessenberg reduction behind StarNEig's C interface.

The product is ``lib/libstarneig.so`` (C ABI, sources in ``csrc/``); this package is the thin host-side
mirror of the reference interface used by the tests and the benchmark.
"""
from .api import *  # noqa: F401,F403
from .api import lib, get_stats, set_profile_level, hessenberg_device, default_panel_width  # noqa: F401
from .api import hessenberg_stage, stage_fetch, starneig_b200_SEP_SM_Reduce  # noqa: F401
from ._lib import Chain, SCHUR_FN, SELECT_FN, REORDER_FN, PREDICATE_FN  # noqa: F401
