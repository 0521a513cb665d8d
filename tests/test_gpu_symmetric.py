"""GPU parity tests of the Newton's-third-law Stokeslet kernel (sym_kernels.cuh): used when the sources are the
leading targets (fiber -> fiber block of apply_matvec).  Same 1e-12 gate as the plain kernel."""
import os

import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu
SL, DL = skb.KERNEL_STOKESLET, skb.KERNEL_STRESSLET
TOL = 1e-12


def _check(u, ref):
    assert np.isfinite(u).all()
    assert rel_max(u, ref) < TOL, rel_max(u, ref)
    assert rel_l2(u, ref) < TOL, rel_l2(u, ref)


@pytest.mark.parametrize("n_src,n_extra", [(1024, 0), (1025, 0), (2048, 300), (3000, 1), (5000, 777), (4097, 4097)])
def test_symmetric_matches_oracle(n_src, n_extra):
    rng = np.random.default_rng(n_src + n_extra)
    rs = rng.uniform(-1, 1, (n_src, 3))
    rt = np.concatenate([rs, rng.uniform(-1, 1, (n_extra, 3))])
    f = rng.uniform(-1, 1, (n_src, 3))
    ref = orc.stokeslet_direct_cpu(rs, f, rt, 1.0)
    with skb.Context(1) as c:
        c.set_symmetric(1)
        c.set_targets(rt)
        c.set_sources(SL, rs)
        u = c.eval(SL, f)
        st = c.stats()
        u_again = c.eval(SL, f)
        c.set_symmetric(0)
        u_plain = c.eval(SL, f)
    _check(u, ref)
    _check(u_plain, ref)
    assert np.array_equal(u, u_again), "symmetric path must be bitwise reproducible"
    assert st["n_pairs"] == n_src * (n_src + n_extra)


def test_symmetric_with_duplicate_nodes_and_zero_strengths():
    rng = np.random.default_rng(5)
    rs = np.repeat(rng.uniform(-1, 1, (600, 3)), 2, axis=0)   # every node twice: r == 0 pairs off the diagonal
    f = rng.uniform(-1, 1, (1200, 3))
    f[::5] = 0.0
    with skb.Context(1) as c:
        c.set_symmetric(1)
        c.set_targets(rs)
        c.set_sources(SL, rs)
        u = c.eval(SL, f)
    _check(u, orc.stokeslet_direct_cpu(rs, f, rs, 1.0))


def test_symmetric_not_used_when_targets_differ():
    rng = np.random.default_rng(6)
    rs = rng.uniform(-1, 1, (2048, 3))
    rt = rs.copy()
    rt[1000, 1] = np.nextafter(rt[1000, 1], 2.0)   # one ulp off: no longer a self-interaction
    f = rng.uniform(-1, 1, (2048, 3))
    with skb.Context(1) as c:
        c.set_symmetric(1)
        c.set_targets(rt)
        c.set_sources(SL, rs)
        u = c.eval(SL, f)
    _check(u, orc.stokeslet_direct_cpu(rs, f, rt, 1.0))


def test_symmetric_accumulate_and_position_update():
    rng = np.random.default_rng(7)
    rs = rng.uniform(-1, 1, (2500, 3))
    rt = np.concatenate([rs, rng.uniform(-1, 1, (100, 3))])
    f, f9 = rng.uniform(-1, 1, (2500, 3)), rng.uniform(-1, 1, (2500, 9))
    with skb.Context(1) as c:
        c.set_symmetric(1)
        c.set_targets(rt)
        c.set_sources(SL, rs)
        c.set_sources(DL, rs)
        out = c.eval(DL, f9)
        c.eval(SL, f, out=out, accumulate=True)
        _check(out, orc.stresslet_direct_cpu(rs, f9, rt, 1.0) + orc.stokeslet_direct_cpu(rs, f, rt, 1.0))
        rs2 = rs + 0.01
        rt2 = np.concatenate([rs2, rt[2500:]])
        c.set_sources(SL, rs2)
        c.set_targets(rt2)
        _check(c.eval(SL, f), orc.stokeslet_direct_cpu(rs2, f, rt2, 1.0))


def test_fused_matvec_flow_with_symmetric_kernel(monkeypatch):
    from test_gpu_flow import ft_of, load, make_system
    monkeypatch.setenv("SKB_SYMMETRIC", "1")
    fib, shell, body = make_system(31, 80, 700, 400, 1, nodes=(16, 24, 32))
    eta = 1.2
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    _check(v, orc.matvec_flow(fib, shell, body, eta))


def test_auto_mode_large_self_interaction():
    # auto: >= 4096 leading-target sources on a single GPU -> symmetric kernel; checked on a target subset
    rng = np.random.default_rng(8)
    n = 20000
    rs = rng.uniform(-3, 3, (n, 3))
    rt = np.concatenate([rs, rng.uniform(-3, 3, (1500, 3))])
    f = rng.uniform(-1, 1, (n, 3))
    with skb.Context(1) as c:
        c.set_targets(rt)
        c.set_sources(SL, rs)
        u = c.eval(SL, f)
    sub = rng.choice(rt.shape[0], 300, replace=False)
    _check(u[sub], orc.stokeslet_direct_cpu(rs, f, rt[sub], 1.0))


@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_row_partition_parts_sum_to_the_full_result(n_parts):
    # one rank per GPU with the symmetric kernel: every part evaluates its block rows; the leading n_src rows are
    # partial sums that add up to the full self-interaction; remainder targets are complete per part
    rng = np.random.default_rng(40 + n_parts)
    n_src = 6000
    rs = rng.uniform(-2, 2, (n_src, 3))
    f = rng.uniform(-1, 1, (n_src, 3))
    extra = rng.uniform(-2, 2, (n_parts * 50, 3))
    total = np.zeros((n_src, 3))
    for p in range(n_parts):
        mine = extra[p * 50:(p + 1) * 50]
        with skb.Context(1) as c:
            c.set_symmetric(1)
            c.set_sym_partition(p, n_parts)
            c.set_targets(np.concatenate([rs, mine]))
            c.set_sources(SL, rs)
            u = c.eval(SL, f)
            assert c.last_eval_was_symmetric()
        total += u[:n_src]
        _check(u[n_src:], orc.stokeslet_direct_cpu(rs, f, mine, 1.0))
    _check(total, orc.stokeslet_direct_cpu(rs, f, rs, 1.0))
