"""skellysim_b200 -- B200-native (sm_100a) backend for SkellySim's hydrodynamic pair-kernel hot path.

The product is the C-ABI shared library ``skellysim_b200/lib/libskelly_b200.so`` (sources in
``skellysim_b200/csrc``, interface in ``include/skelly_b200.h``).  This package is only the thin
ctypes binding used by the tests and bench.py.  There is no CPU fallback: importing works without a
GPU (so the C-ABI export test can run), every compute call needs a B200.
"""
from .capi import (DENSE_M_INV, DENSE_STRESSLET_PLUS_COMPLEMENTARY, KERNEL_STOKESLET, KERNEL_STRESSLET, Context, Dense,
                   Flow, MultiFlow, SkbError, build_library, library, library_path,
                   stokeslet_direct, stresslet_direct)

__all__ = ["DENSE_M_INV", "DENSE_STRESSLET_PLUS_COMPLEMENTARY", "KERNEL_STOKESLET", "KERNEL_STRESSLET", "Context",
           "Dense", "Flow", "MultiFlow", "SkbError", "build_library", "library",
           "library_path", "stokeslet_direct", "stresslet_direct"]
