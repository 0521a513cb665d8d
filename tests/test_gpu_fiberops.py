"""GPU parity tests of the per-fiber dense operators (include/skelly_b200_flow.h
skb_flow_apply_fiber_force / skb_flow_fiber_matvec / skb_flow_apply_matvec) against the oracle's restatement of
FiberContainerFiniteDifference::apply_fiber_force (fcfd.cpp:272-287), FiberFiniteDifference::matvec
(fiber_finite_difference.cpp:276-312) and System::apply_matvec (system.cpp:298-318).
Tolerance 1e-12 relative (max-norm and l2): FP64 dot products of <= 4*64 terms in a different summation order."""
import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_l2, rel_max
from test_gpu_flow import ft_of, load, make_system

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _check(u, ref, tol=TOL):
    assert np.isfinite(u).all()
    assert rel_max(u, ref) < tol, rel_max(u, ref)
    assert rel_l2(u, ref) < tol, rel_l2(u, ref)


def make_ops(fib, seed):
    """Random dense stand-ins with the shapes of the reference's members (the kernels are plain linear algebra)."""
    rng = np.random.default_rng(seed)
    n_nodes = [int(n) for n in fib["n_nodes"]]
    ops = dict(n_nodes=n_nodes, A=[], force=[], D_1_0={}, P={}, length_prev=[], plus=[])
    xs = []
    for n in n_nodes:
        ops["A"].append(rng.normal(size=(4 * n, 4 * n)) / np.sqrt(4 * n))
        ops["force"].append(rng.normal(size=(3 * n, 4 * n)) / np.sqrt(4 * n))
        t = rng.normal(size=(n, 3))
        xs.append(t / np.linalg.norm(t, axis=1)[:, None])
        ops["length_prev"].append(rng.uniform(0.5, 2.0))
        ops["plus"].append(int(rng.integers(0, 2)))
        if n not in ops["D_1_0"]:
            ops["D_1_0"][n] = rng.normal(size=(n, n))
            ops["P"][n] = rng.normal(size=(4 * n - 14, 4 * n)) / np.sqrt(4 * n)
    ops["xs"] = np.concatenate(xs) if xs else np.zeros((0, 3))
    return ops


def load_ops(fl, ops):
    for n in ops["D_1_0"]:
        fl.set_fiber_class(n, ops["D_1_0"][n], ops["P"][n])
    fl.set_fiber_operators(ops["A"], ops["force"], ops["xs"], ops["length_prev"], ops["plus"])


NODES = (4, 5, 8, 16, 17, 32, 48, 64)


def test_apply_fiber_force_ragged():
    fib, shell, body = make_system(21, 57, 0, 0, 0, nodes=NODES)
    ops = make_ops(fib, 1)
    x = np.random.default_rng(2).normal(size=4 * fib["pos"].shape[0])
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        fw = fl.apply_fiber_force(x)
        assert fl.stats()["launches"] == 1
    _check(fw, orc.apply_fiber_force(ops["force"], x, ops["n_nodes"]))


@pytest.mark.parametrize("with_boundary", [False, True])
def test_fiber_matvec_ragged(with_boundary):
    fib, shell, body = make_system(22, 41, 0, 0, 0, nodes=NODES)
    ops = make_ops(fib, 3)
    rng = np.random.default_rng(4)
    nf = fib["pos"].shape[0]
    x, v = rng.normal(size=4 * nf), rng.normal(size=(nf, 3))
    vb = rng.normal(size=(len(ops["n_nodes"]), 7)) if with_boundary else None
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        res = fl.fiber_matvec(x, v, vb)
        res2 = fl.fiber_matvec(x, v, vb)
    assert np.array_equal(res, res2)  # fixed summation order
    _check(res, orc.fiber_container_matvec(ops, x, v, vb))


def test_fiber_matvec_long_fiber():
    """One fiber of 300 nodes: 1200 columns, several row blocks, shared memory above the smallest sizes."""
    fib, shell, body = make_system(23, 3, 0, 0, 0, nodes=(300,))
    ops = make_ops(fib, 5)
    rng = np.random.default_rng(6)
    nf = fib["pos"].shape[0]
    x, v = rng.normal(size=4 * nf), rng.normal(size=(nf, 3))
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        fw = fl.apply_fiber_force(x)
        res = fl.fiber_matvec(x, v, None)
    _check(fw, orc.apply_fiber_force(ops["force"], x, ops["n_nodes"]))
    _check(res, orc.fiber_container_matvec(ops, x, v, None))


@pytest.mark.parametrize("shape", [(60, 700, 400, 2), (25, 0, 0, 0), (30, 500, 0, 0)])
def test_apply_matvec_end_to_end(shape):
    fib, shell, body = make_system(24, *shape, nodes=(8, 16, 32, 64))
    ops = make_ops(fib, 7)
    rng = np.random.default_rng(8)
    nf, ns = fib["pos"].shape[0], shell["pos"].shape[0]
    x = rng.normal(size=4 * nf)
    link = rng.normal(size=(len(ops["n_nodes"]), 7))
    eta = 0.8
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        res, v_s, v_b = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, link)
        # the two-step host path gives the same numbers
        fw = fl.apply_fiber_force(x)
        v_all = fl.matvec(fw, shell["density"], body["density"], ft_of(body), eta)
        res_2 = fl.fiber_matvec(x, v_all[:nf], link)
    ref_res, ref_v = orc.apply_matvec_fibers(fib, shell, body, ops, x, eta, link)
    _check(res, ref_res)
    if ns:
        _check(v_s, ref_v[nf:nf + ns])
    if body["pos"].shape[0]:
        _check(v_b, ref_v[nf + ns:])
    assert np.array_equal(res, res_2)
    assert np.array_equal(v_s, v_all[nf:nf + ns])


@pytest.mark.parametrize("seed,clamp,pin", [(1, 0.0, 0.0), (2, 0.5, 0.5)])
def test_apply_matvec_with_reference_shaped_operators(seed, clamp, pin):
    """The same path on operators built the way the reference builds them (oracle/fiber_fd.py: Fornberg derivative
    matrices, barycentric down-sampling, update_linear_operator / apply_bc_rectangular / update_force_operator):
    P_downsample_bc is block diagonal (the kernel walks each row's non-zero columns only), D_1_0 banded, A_ carries its
    14 boundary rows.  Allowed node counts only (fiber_finite_difference.cpp:522)."""
    from oracle import fiber_fd
    fib, shell, body = make_system(40 + seed, 45, 500, 260, 2, nodes=(8, 16, 24, 32, 48, 64, 96))
    ops = fiber_fd.suspension_operators(fib["pos"], fib["n_nodes"], fib["lengths"], eta=0.9, seed=seed,
                                        clamp_fraction=clamp, pin_fraction=pin)
    for n, P in ops["P"].items():
        assert np.count_nonzero(P) <= (4 * n - 14) * n        # block diagonal: at most n non-zeros per row
    rng = np.random.default_rng(seed)
    nf, ns = fib["pos"].shape[0], shell["pos"].shape[0]
    x = rng.normal(size=4 * nf)
    link = rng.normal(size=(len(ops["n_nodes"]), 7))
    eta = 0.9
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        fl.set_fiber_preconditioner([np.linalg.inv(A) for A in ops["A"]])
        res, v_s, v_b = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, link)
        fw = fl.apply_fiber_force(x)
        y = fl.apply_fiber_preconditioner(x)
    ref_res, ref_v = orc.apply_matvec_fibers(fib, shell, body, ops, x, eta, link)
    _check(fw, orc.apply_fiber_force(ops["force"], x, ops["n_nodes"]))
    _check(res, ref_res)
    _check(v_s, ref_v[nf:nf + ns])
    _check(v_b, ref_v[nf + ns:])
    # fc.apply_preconditioner: LU solve in the reference, explicit inverse here; A_ of a discretised 4th-order
    # operator is ill conditioned, so compare through the residual A y = x instead of y itself
    off = 0
    for A, n in zip(ops["A"], ops["n_nodes"]):
        r = A @ y[4 * off:4 * off + 4 * n] - x[4 * off:4 * off + 4 * n]
        assert np.abs(r).max() <= 1e-6 * max(1.0, np.abs(A).max() * np.abs(y[4 * off:4 * off + 4 * n]).max())
        off += n


def test_operator_errors():
    fib, shell, body = make_system(25, 6, 50, 0, 0, nodes=(8, 16))
    ops = make_ops(fib, 9)
    x = np.zeros(4 * fib["pos"].shape[0])
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        with pytest.raises(skb.SkbError, match="set_fiber_operators"):
            fl.apply_fiber_force(x)
        # a node count without class matrices
        only = sorted(ops["D_1_0"])[0]
        fl.set_fiber_class(only, ops["D_1_0"][only], ops["P"][only])
        if len(ops["D_1_0"]) > 1:
            with pytest.raises(skb.SkbError, match="set_fiber_class"):
                fl.set_fiber_operators(ops["A"], ops["force"], ops["xs"], ops["length_prev"], ops["plus"])
        load_ops(fl, ops)
        fl.apply_fiber_force(x)
        # new fibers invalidate the operators
        load(fl, fib, shell, body)
        with pytest.raises(skb.SkbError, match="set_fiber_operators"):
            fl.fiber_matvec(x, np.zeros((fib["pos"].shape[0], 3)))
        load_ops(fl, ops)
        # a window drops the resident operators (they belong to the fiber rows of the target list) ...
        n0 = ops["n_nodes"][0]
        fl.set_target_window(0, n0)
        with pytest.raises(skb.SkbError, match="set_fiber_operators"):
            fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), 1.0)
        # ... and the one-call host form needs every row
        fl.set_fiber_operators(ops["A"][:1], ops["force"][:1], ops["xs"][:n0], ops["length_prev"][:1], ops["plus"][:1])
        with pytest.raises(skb.SkbError, match="full target window"):
            fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), 1.0)
        with pytest.raises((skb.SkbError, ValueError)):
            fl.set_fiber_class(3, np.zeros((3, 3)), np.zeros((0, 12)))


def test_no_fibers():
    fib, shell, body = make_system(26, 0, 300, 200, 1)
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        fl.set_fiber_operators([], [], np.zeros((0, 3)), [], [])
        res, v_s, v_b = fl.apply_matvec(np.zeros(0), shell["density"], body["density"], ft_of(body), 1.0)
        v_all = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), 1.0)
    assert res.shape == (0,)
    assert np.array_equal(v_s, v_all[:300]) and np.array_equal(v_b, v_all[300:])


def test_apply_matvec_with_periphery_dense_operator():
    """res_shell = stresslet_plus_complementary_ * x_shell + v_shell (periphery.cpp:38-47) formed on the device.  With
    the overlap off (one GEMV kernel at the end of the stream) it is bit-identical to skb_dense_apply on the v_shell that
    skb_flow_apply_matvec returns; the default (background row streamer beside the pair kernels) sums a row in another
    order: same result to rounding, identical from call to call."""
    fib, shell, body = make_system(27, 40, 600, 300, 1, nodes=(8, 16, 32))
    ops = make_ops(fib, 11)
    rng = np.random.default_rng(12)
    nf, ns = fib["pos"].shape[0], shell["pos"].shape[0]
    x = rng.normal(size=4 * nf)
    M = rng.normal(size=(3 * ns, 3 * ns)) / np.sqrt(3 * ns)
    eta = 1.1
    with skb.Flow(0) as fl, skb.Dense(1) as dn:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, M)
        res, v_s, v_b = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta)
        res_o, res_shell_o, v_b_o = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, dense=dn)
        res_o2, res_shell_o2, _ = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, dense=dn)
        fl.set_overlap(False)
        res_d, res_shell, v_b_d = fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, dense=dn)
        fl.set_overlap(True)
        two_step = dn.apply(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, shell["density"].reshape(-1), v_s.reshape(-1))
        # wrong size is refused
        dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, M[:-3, :-3])
        with pytest.raises(skb.SkbError, match="own rows"):
            fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta, dense=dn)
    assert np.array_equal(res, res_d) and np.array_equal(v_b, v_b_d)
    assert np.array_equal(res, res_o) and np.array_equal(v_b, v_b_o)
    assert np.array_equal(res_shell.reshape(-1), two_step)
    assert np.array_equal(res_shell_o, res_shell_o2)
    _check(res_shell_o.reshape(-1), orc.periphery_dense_apply(M, shell["density"].reshape(-1), v_s.reshape(-1)))
    _check(res_shell.reshape(-1), orc.periphery_dense_apply(M, shell["density"].reshape(-1), v_s.reshape(-1)))


# ---- one rank per GPU, simulated rank by rank on one device: the reference's MPI decomposition as target ranges ----

def _own_ops(ops, off, f0, f1):
    return dict(ops, n_nodes=ops["n_nodes"][f0:f1], A=ops["A"][f0:f1], force=ops["force"][f0:f1],
                xs=ops["xs"][off[f0]:off[f1]], length_prev=ops["length_prev"][f0:f1], plus=ops["plus"][f0:f1])


@pytest.mark.parametrize("world", [2, 3])
def test_target_ranges_rank_decomposition(world):
    from skellysim_b200.distributed import reference_rank_ranges
    fib, shell, body = make_system(28, 23, 500, 240, 2, nodes=(8, 16, 32))
    ops = make_ops(fib, 13)
    rng = np.random.default_rng(14)
    n_nodes = ops["n_nodes"]
    off = np.concatenate([[0], np.cumsum(n_nodes)])
    nf, ns, nb = int(off[-1]), shell["pos"].shape[0], body["pos"].shape[0]
    x = rng.normal(size=4 * nf)
    link = rng.normal(size=(len(n_nodes), 7))
    eta = 1.2
    ref_res, ref_v = orc.apply_matvec_fibers(fib, shell, body, ops, x, eta, link)
    ref_fw = orc.apply_fiber_force(ops["force"], x, n_nodes)
    res_parts, v_f, v_s, v_b = [], [], [], []
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        for n in ops["D_1_0"]:
            fl.set_fiber_class(n, ops["D_1_0"][n], ops["P"][n])
        fw_parts = []
        for rank in range(world):
            f0, f1, s0, s1, b0, b1 = reference_rank_ranges(len(n_nodes), ns, nb, rank, world)
            fl.set_target_ranges(f0, f1, s0, s1, b0, b1)
            own = _own_ops(ops, off, f0, f1)
            fl.set_fiber_operators(own["A"], own["force"], own["xs"], own["length_prev"], own["plus"])
            fw_parts.append(fl.apply_fiber_force(x[4 * off[f0]:4 * off[f1]]))
        fw_all = np.concatenate(fw_parts)            # the all-gather of the ranks
        _check(fw_all, ref_fw)
        for rank in range(world):
            f0, f1, s0, s1, b0, b1 = reference_rank_ranges(len(n_nodes), ns, nb, rank, world)
            fl.set_target_ranges(f0, f1, s0, s1, b0, b1)
            own = _own_ops(ops, off, f0, f1)
            fl.set_fiber_operators(own["A"], own["force"], own["xs"], own["length_prev"], own["plus"])
            v_own = fl.matvec(fw_all, shell["density"], body["density"], ft_of(body), eta)
            n_own = int(off[f1] - off[f0])
            assert v_own.shape[0] == n_own + (s1 - s0) + (b1 - b0)
            res_parts.append(fl.fiber_matvec(x[4 * off[f0]:4 * off[f1]], v_own[:n_own], link[f0:f1]))
            v_f.append(v_own[:n_own])
            v_s.append(v_own[n_own:n_own + s1 - s0])
            v_b.append(v_own[n_own + s1 - s0:])
            with pytest.raises(skb.SkbError, match="full target window"):
                fl.apply_matvec(x, shell["density"], body["density"], ft_of(body), eta)
    _check(np.concatenate(v_f + v_s + v_b), ref_v)
    _check(np.concatenate(res_parts), ref_res)


def test_fiber_operator_device_pointer_forms():
    torch = pytest.importorskip("torch")
    fib, shell, body = make_system(29, 19, 0, 0, 0, nodes=(8, 16, 17))
    ops = make_ops(fib, 15)
    rng = np.random.default_rng(16)
    off = np.concatenate([[0], np.cumsum(ops["n_nodes"])])
    f0, f1 = 4, 15
    own = _own_ops(ops, off, f0, f1)
    n_own = int(off[f1] - off[f0])
    x, v, link = rng.normal(size=4 * n_own), rng.normal(size=(n_own, 3)), rng.normal(size=(f1 - f0, 7))
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_x, d_v, d_link = t(x), t(v), t(link)
    d_fw = torch.empty((n_own, 3), dtype=torch.float64, device=dev)
    d_res = torch.empty(4 * n_own, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        for n in ops["D_1_0"]:
            fl.set_fiber_class(n, ops["D_1_0"][n], ops["P"][n])
        fl.set_target_ranges(f0, f1, 0, 0, 0, 0)
        fl.set_fiber_operators(own["A"], own["force"], own["xs"], own["length_prev"], own["plus"])
        fl.apply_fiber_force_device(d_x.data_ptr(), d_fw.data_ptr(), stream)
        fl.fiber_matvec_device(d_x.data_ptr(), d_v.data_ptr(), d_link.data_ptr(), d_res.data_ptr(), stream)
        torch.cuda.synchronize()
        fw_h = fl.apply_fiber_force(x)
        res_h = fl.fiber_matvec(x, v, link)
        # a window that cuts a fiber cannot hold per-fiber operators
        fl.set_target_window(3, int(off[-1]))
        with pytest.raises(skb.SkbError, match="cuts a fiber"):
            fl.set_fiber_operators(ops["A"], ops["force"], ops["xs"][3:], ops["length_prev"], ops["plus"])
    assert np.array_equal(d_fw.cpu().numpy(), fw_h) and np.array_equal(d_res.cpu().numpy(), res_h)
    _check(fw_h, orc.apply_fiber_force(own["force"], x, own["n_nodes"]))
    _check(res_h, orc.fiber_container_matvec(own, x, v, link))


def test_fiber_preconditioner():
    """fc.apply_preconditioner (fcfd.cpp:331-339): the per-fiber LU solve as a GEMV over the explicit inverse.
    Well-conditioned A_ (I + small perturbation) so that inverse-vs-substitution rounding stays below the gate."""
    fib, shell, body = make_system(30, 33, 0, 0, 0, nodes=NODES)
    ops = make_ops(fib, 17)
    rng = np.random.default_rng(18)
    ops["A"] = [np.eye(4 * n) + 0.3 * rng.normal(size=(4 * n, 4 * n)) / np.sqrt(4 * n) for n in ops["n_nodes"]]
    A_inv = [np.linalg.inv(a) for a in ops["A"]]
    nf = fib["pos"].shape[0]
    x = rng.normal(size=4 * nf)
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        for n in ops["D_1_0"]:
            fl.set_fiber_class(n, ops["D_1_0"][n], ops["P"][n])
        with pytest.raises(skb.SkbError, match="set_fiber_operators"):
            fl.set_fiber_preconditioner(A_inv)
        fl.set_fiber_operators(ops["A"], ops["force"], ops["xs"], ops["length_prev"], ops["plus"])
        with pytest.raises(skb.SkbError, match="set_fiber_preconditioner"):
            fl.apply_fiber_preconditioner(x)
        fl.set_fiber_preconditioner(A_inv)
        y = fl.apply_fiber_preconditioner(x)
        assert fl.stats()["launches"] == 1
        # P^-1 (A x) = x through the two device operators
        ax = fl.fiber_matvec(x, np.zeros((nf, 3)))        # v = 0, no link conditions: res = A_ x
        back = fl.apply_fiber_preconditioner(ax)
        # new operators invalidate the preconditioner
        fl.set_fiber_operators(ops["A"], ops["force"], ops["xs"], ops["length_prev"], ops["plus"])
        with pytest.raises(skb.SkbError, match="set_fiber_preconditioner"):
            fl.apply_fiber_preconditioner(x)
    _check(y, orc.fiber_apply_preconditioner(ops["A"], x, ops["n_nodes"]), tol=1e-12)
    _check(back, x, tol=1e-12)
