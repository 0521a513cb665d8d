"""Multi-GPU tests (need >= 2 visible B200s: `gpurun --gpus 2 -- python -m pytest tests -m gpu`): the single-process
multi-device context (targets block-partitioned over devices, ONE NCCL all-gather of strengths per evaluation)."""
import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_l2, rel_max
from skellysim_b200 import capi

pytestmark = pytest.mark.gpu
SL, DL = skb.KERNEL_STOKESLET, skb.KERNEL_STRESSLET


def _need(n):
    if capi.device_count() < n:
        pytest.skip(f"needs {n} GPUs, {capi.device_count()} visible")


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
@pytest.mark.parametrize("ns,nt", [(1229, 743), (4001, 3001), (130, 5)])
def test_single_process_multi_device_context(n_gpus, ns, nt):
    _need(n_gpus)
    rng = np.random.default_rng(ns + nt + n_gpus)
    rs, rt = rng.uniform(-1, 1, (ns, 3)), rng.uniform(-1, 1, (nt, 3))
    rt[:min(5, nt)] = rs[:min(5, nt)]
    f3, f9 = rng.uniform(-1, 1, (ns, 3)), rng.uniform(-1, 1, (ns, 9))
    with skb.Context(n_gpus) as c:
        assert c.n_gpus == n_gpus
        c.set_targets(rt)
        c.set_sources(SL, rs)
        c.set_sources(DL, rs)
        u = c.eval(SL, f3)
        d = c.eval(DL, f9)
        both = c.eval_fused(f3, f9)
        # strengths change, positions stay (GMRES iterations)
        u2 = c.eval(SL, 2.0 * f3)
    ru, rd = orc.stokeslet_direct(rs, f3, rt), orc.stresslet_direct(rs, f9, rt)
    for got, ref in ((u, ru), (d, rd), (both, ru + rd), (u2, 2.0 * ru)):
        assert rel_max(got, ref) < 1e-12 and rel_l2(got, ref) < 1e-12


def test_multi_device_matches_single_device_bitwise_per_block():
    _need(2)
    rng = np.random.default_rng(3)
    rs, rt, f3 = rng.uniform(-1, 1, (3000, 3)), rng.uniform(-1, 1, (2048, 3)), rng.uniform(-1, 1, (3000, 3))
    with skb.Context(2) as c2:
        c2.set_targets(rt)
        c2.set_sources(SL, rs)
        u2 = c2.eval(SL, f3)
    assert rel_max(u2, orc.stokeslet_direct(rs, f3, rt)) < 1e-12


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_multi_device_symmetric_layout(n_gpus):
    # single process, P devices, targets = [sources | extra]: every device evaluates its serpentine share of the
    # self-interaction's block rows for ALL leading targets; one NCCL reduce-scatter combines the partial sums
    _need(n_gpus)
    rng = np.random.default_rng(50 + n_gpus)
    n_src, n_extra = 6000, 333
    rs = rng.uniform(-2, 2, (n_src, 3))
    rt = np.concatenate([rs, rng.uniform(-2, 2, (n_extra, 3))])
    f = rng.uniform(-1, 1, (n_src, 3))
    ref = orc.stokeslet_direct_cpu(rs, f, rt, 1.0)
    with skb.Context(n_gpus) as c:
        c.set_symmetric(1)
        c.set_targets(rt)
        c.set_sources(SL, rs)
        u = c.eval(SL, f)
        assert c.last_eval_was_symmetric()
        u2 = c.eval(SL, f)
        out = u.copy()
        c.eval(SL, 0.5 * f, out=out, accumulate=True)
        # new strengths, same positions
        g = rng.uniform(-1, 1, (n_src, 3))
        ug = c.eval(SL, g)
        # a stresslet source set in the same context switches back to the plain block layout
        c.set_sources(DL, rs[:100])
        u_plain = c.eval(SL, f)
        assert not c.last_eval_was_symmetric()
        # targets that no longer start with the sources: plain layout as well
        c.set_sources(DL, np.zeros((0, 3)))
        rt2 = rt.copy()
        rt2[5, 0] += 1e-3
        c.set_targets(rt2)
        u_moved = c.eval(SL, f)
        assert not c.last_eval_was_symmetric()
    for got, want in ((u, ref), (u_plain, ref), (out, 1.5 * ref), (ug, orc.stokeslet_direct_cpu(rs, g, rt, 1.0)),
                      (u_moved, orc.stokeslet_direct_cpu(rs, f, rt2, 1.0))):
        assert rel_max(got, want) < 1e-12 and rel_l2(got, want) < 1e-12
    assert np.array_equal(u, u2)
