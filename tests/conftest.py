import os
import sys

import numpy as np
import pytest

try:  # torch first: it must load ITS bundled libnccl.so.2 before the library dlopens "libnccl.so.2" for a multi-device
    import torch  # noqa: F401  context (a system copy loaded earlier would be reused by torch and may be older)
except Exception:  # pragma: no cover
    torch = None

# two group members on ONE GPU (tests/test_gpu_mflow.py) spin-wait on each other's flags from different streams: give
# every stream its own hardware queue, so that no launch is serialised behind a waiting kernel (read at CUDA init)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import skellysim_b200.capi as capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: a silent skip would read as green.
    # A plain `pytest` (no -m) on a box without a GPU skips the gpu-marked tests instead of failing them.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (run `pytest -m gpu` on the GPU box; that invocation never skips)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cases():
    import glob
    out = {}
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "ref_numba_*.npz"))):
        name = os.path.basename(p)[len("ref_numba_"):-4]
        out[name] = dict(np.load(p))
    assert out, "golden fixtures missing"
    return out


def rel_max(a, b):
    """max|a-b| / max|b|  (the parity metric of SURVEY.md 8d / BASELINE.md 2.4)"""
    return float(np.abs(a - b).max() / np.abs(b).max())


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2  (performance_hydrodynamics_combined.cpp:95)"""
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
