"""GPU parity tests of the opt-in fused self-exclusion: the pair kernels skip every pair whose two
nodes carry the same id (fiber index) instead of computing all pairs and subtracting each fiber's own block
(fiber_container_finite_difference.cpp:203-210).  Checked against "all pairs - same-id pairs" of the scalar oracle and
against the reference-shaped compute-then-subtract flow.  Tolerance <= 1e-12 relative."""
import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_l2, rel_max
from test_gpu_flow import ft_of, load, make_system

pytestmark = pytest.mark.gpu
SL = skb.KERNEL_STOKESLET
TOL = 1e-12


def _check(u, ref, tol=TOL):
    assert np.isfinite(u).all()
    assert rel_max(u, ref) < tol, rel_max(u, ref)
    assert rel_l2(u, ref) < tol, rel_l2(u, ref)


def _ragged_ids(rng, n, sizes=(8, 16, 24, 32, 48, 64, 96, 128)):
    ids, g = [], 0
    while len(ids) < n:
        ids += [g] * int(rng.choice(sizes))
        g += 1
    return np.asarray(ids[:n], dtype=np.int32)


@pytest.mark.parametrize("n_src,n_extra,sym", [(700, 0, 0), (700, 333, 0), (3000, 1, 1), (5000, 777, 1),
                                               (4097, 4097, 1), (9000, 0, 1), (9000, 100, 0)])
def test_exclusion_matches_all_pairs_minus_same_id_pairs(n_src, n_extra, sym):
    rng = np.random.default_rng(n_src + 3 * n_extra + sym)
    rs = rng.uniform(-1, 1, (n_src, 3))
    rt = np.concatenate([rs, rng.uniform(-1, 1, (n_extra, 3))])
    f = rng.uniform(-1, 1, (n_src, 3))
    ids = _ragged_ids(rng, n_src)
    ref = orc.stokeslet_direct_excluding(rs, f, rt, ids)
    with skb.Context(1) as c:
        c.set_symmetric(sym)
        c.set_targets(rt)
        c.set_sources(SL, rs)
        c.set_source_exclusion_ids(ids)
        u = c.eval(SL, f)
        was_sym = c.last_eval_was_symmetric()
        u2 = c.eval(SL, f)
        c.set_source_exclusion_ids(None)
        u_all = c.eval(SL, f)
    assert was_sym == bool(sym)
    _check(u, ref)
    assert np.array_equal(u, u2), "bitwise reproducible"
    _check(u_all, orc.stokeslet_direct(rs, f, rt))      # switching it off restores the plain sum


def test_exclusion_needs_the_sources_as_leading_targets():
    rng = np.random.default_rng(1)
    rs = rng.uniform(-1, 1, (600, 3))
    with skb.Context(1) as c:
        c.set_targets(rng.uniform(-1, 1, (600, 3)))
        c.set_sources(SL, rs)
        c.set_source_exclusion_ids(np.arange(600) // 32)
        with pytest.raises(skb.SkbError):
            c.eval(SL, rng.uniform(-1, 1, (600, 3)))
        with pytest.raises(skb.SkbError):
            c.set_source_exclusion_ids(np.arange(599))  # wrong count


@pytest.mark.parametrize("n_fibers,force_sym", [(37, 0), (37, 1), (160, 1)])
def test_matvec_fused_self_exclusion_equals_compute_then_subtract(n_fibers, force_sym, monkeypatch):
    # the reference's semantics (oracle.matvec_flow: all pairs, then the regularised self block is subtracted) and the
    # fused kernels agree to rounding: no two nodes of one fiber are closer than the 1e-5 regularisation threshold
    if force_sym:
        monkeypatch.setenv("SKB_SYMMETRIC", "1")
    fib, shell, body = make_system(11 + n_fibers, n_fibers, 500, 300, 2, nodes=(16, 32, 48, 64, 96))
    eta = 0.7
    ref = orc.matvec_flow(fib, shell, body, eta)
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v_sub = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        fl.set_self_exclusion(True)
        v_fused = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        pairs_fused = fl.stats()["n_pairs"]
        fl.set_self_exclusion(False)
        v_sub2 = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    _check(v_sub, ref)
    _check(v_fused, ref)
    assert np.array_equal(v_sub, v_sub2)
    assert pairs_fused > 0
