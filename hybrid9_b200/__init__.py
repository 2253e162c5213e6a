"""hybrid9_b200 -- B200-native drop-in for HYBRID9's per-grid-cell time-stepping path.

The product is ``libh9gpu.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/h9gpu.h``).  This package is the thin Python host layer used by the
tests and the benchmark; it mirrors the calls the Fortran host makes
(INTEGRATION.md).  There is no CPU implementation of the physics here: loading
fails loudly when the CUDA library is missing.
"""
from .host import H9, H9Error, H9Fault, MATH_EXACT, MATH_FAST, load_library  # noqa: F401
from .state import H9State, init_state  # noqa: F401
from .calendar import time_boy, decade_days, year_index_of_days  # noqa: F401
from .spinup import load_restart, save_restart, spin_up  # noqa: F401

__all__ = [
    "H9", "H9Error", "H9Fault", "H9State", "init_state", "MATH_EXACT", "MATH_FAST",
    "load_library", "time_boy", "decade_days", "year_index_of_days", "save_restart",
    "load_restart", "spin_up",
]
