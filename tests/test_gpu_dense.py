"""GPU parity tests of the periphery dense operators (include/skelly_b200_dense.h) against numpy / exactly rounded
row sums (Periphery::matvec, Periphery::apply_preconditioner, periphery.cpp:21-47)."""
import math

import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from skellysim_b200 import capi

pytestmark = pytest.mark.gpu


def _backward_err(y, A, x, v=None):
    """componentwise backward error: |y - (Ax+v)| / (|A||x| + |v|)"""
    ref = orc.periphery_dense_apply(A, x, v)
    scale = np.abs(A) @ np.abs(x) + (np.abs(v) if v is not None else 0.0)
    return float(np.max(np.abs(y - ref) / np.maximum(scale, 1e-300)))


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 5), (300, 300), (1201, 1201), (1800, 1802), (64, 4099), (3000, 3000)])
def test_dense_apply_matches_numpy(rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    A = rng.normal(size=(rows, cols))
    x = rng.normal(size=cols)
    v = rng.normal(size=rows)
    with skb.Dense(1) as dn:
        dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, A)
        dn.set_matrix(skb.DENSE_M_INV, 2.0 * A)
        y = dn.apply(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, x, v)     # Periphery::matvec
        z = dn.apply(skb.DENSE_M_INV, x)                               # Periphery::apply_preconditioner
        y2 = dn.apply(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, x, v)
        st = dn.stats()
    assert _backward_err(y, A, x, v) < 1e-13
    assert _backward_err(z, 2.0 * A, x) < 1e-13
    assert np.array_equal(y, y2)
    assert st["bytes"] == 8 * rows * cols
    # a few rows against exactly rounded sums
    for r in (0, rows // 2, rows - 1):
        exact = math.fsum(A[r] * x) + v[r]
        assert abs(y[r] - exact) <= 1e-13 * (np.abs(A[r]) @ np.abs(x) + abs(v[r]))


def test_dense_multi_gpu_row_blocks():
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(3)
    A = rng.normal(size=(2501, 2500))
    x = rng.normal(size=2500)
    with skb.Dense(2) as dn:
        dn.set_matrix(0, A)
        y = dn.apply(0, x)
    assert _backward_err(y, A, x) < 1e-13


def test_dense_errors():
    with skb.Dense(1) as dn:
        dn.shape[0] = (3, 3)
        with pytest.raises(skb.SkbError):
            dn.apply(0, np.zeros(3))  # matrix never set


@pytest.mark.parametrize("rows,cols", [(1, 2), (3, 2), (7, 6), (5, 254), (4, 256), (9, 258), (301, 300), (1203, 1200),
                                       (130, 4098), (64, 4099), (2999, 3000), (18, 18000)])
def test_background_streamer_matches_numpy(rows, cols):
    """skb_dense_apply_background_device (csrc/stream_kernels.cuh): row groups of 4 with a ragged last group, column
    chunks of 256 with a ragged last chunk, fewer groups than SMs, an odd column count (classic-kernel fallback)."""
    import torch
    rng = np.random.default_rng(rows * 11 + cols)
    A = rng.normal(size=(rows, cols))
    x = rng.normal(size=cols)
    dev = torch.device("cuda", 0)
    d_x = torch.from_numpy(x).to(dev)
    d_y = torch.full((rows,), float("nan"), dtype=torch.float64, device=dev)
    with skb.Dense(1) as dn:
        dn.set_matrix(skb.DENSE_M_INV, A)
        st = torch.cuda.current_stream().cuda_stream
        dn.apply_background_device(skb.DENSE_M_INV, d_x.data_ptr(), d_y.data_ptr(), st)
        torch.cuda.synchronize()
        y = d_y.cpu().numpy().copy()
        d_y.fill_(float("nan"))
        dn.apply_background_device(skb.DENSE_M_INV, d_x.data_ptr(), d_y.data_ptr(), st)
        torch.cuda.synchronize()
        y2 = d_y.cpu().numpy()
    assert _backward_err(y, A, x) < 1e-13
    assert np.array_equal(y, y2)  # ticket order does not change a row's sum
    for r in (0, rows // 2, rows - 1):
        exact = math.fsum(A[r] * x)
        assert abs(y[r] - exact) <= 1e-13 * (np.abs(A[r]) @ np.abs(x))


def test_background_streamer_beside_a_busy_stream():
    """The streamer on a side stream while the main stream is busy with FP64 work: same result, and both finish."""
    import torch
    rng = np.random.default_rng(5)
    rows = cols = 6000
    A = rng.normal(size=(rows, cols))
    x = rng.normal(size=cols)
    dev = torch.device("cuda", 0)
    d_x = torch.from_numpy(x).to(dev)
    d_y = torch.zeros(rows, dtype=torch.float64, device=dev)
    busy = torch.randn(2048, 2048, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream()
    with skb.Dense(1) as dn:
        dn.set_matrix(skb.DENSE_M_INV, A)
        torch.cuda.synchronize()
        for _ in range(3):
            busy2 = busy @ busy
            dn.apply_background_device(skb.DENSE_M_INV, d_x.data_ptr(), d_y.data_ptr(), side.cuda_stream)
        torch.cuda.synchronize()
        y = d_y.cpu().numpy()
    assert torch.isfinite(busy2).all()
    assert _backward_err(y, A, x) < 1e-13
