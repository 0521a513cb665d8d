"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares and
the reference-named C++ symbols, and refuses (loudly, with an error code) to compute without a GPU."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import skellysim_b200 as skb
from skellysim_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names += re.findall(r"SKB_API\s+[\w\s\*]+?\b(skb_\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_something():
    names = _declared_functions()
    assert "skb_eval" in names and "skb_stokeslet_direct" in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    L = C.CDLL(skb.library_path())
    missing = [n for n in _declared_functions() if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_library_exports_reference_named_entry_points():
    # kernels::stokeslet_direct_gpu_impl / stresslet_direct_gpu_impl, include/kernels.hpp:17-20 of the reference
    L = C.CDLL(skb.library_path())
    for sym in ("_ZN7kernels25stokeslet_direct_gpu_implEPKdS1_iS1_Pdi",
                "_ZN7kernels25stresslet_direct_gpu_implEPKdS1_iS1_Pdi"):
        assert hasattr(L, sym), sym


def test_python_binding_covers_header():
    skb.library()
    missing = [n for n in _declared_functions() if n not in capi.BOUND_FUNCTIONS]
    assert not missing, f"declared in include/*.h but without a ctypes signature in capi.py: {missing}"


def test_no_cpu_fallback_without_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(skb.SkbError, match="no CPU fallback"):
        skb.Context(1)
    with pytest.raises(skb.SkbError):
        skb.stokeslet_direct(np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)))


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under skellysim_b200/ or include/ may reference it
    bad = []
    for base in ("skellysim_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"\boracle\b", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
