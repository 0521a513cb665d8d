"""GPU parity tests of the flow() layer (include/skelly_b200_flow.h) against the oracle's restatement of
FiberContainerFiniteDifference::flow, Periphery::flow, BodyContainer::flow and the hydrodynamic part of
System::apply_matvec.  Tolerance <= 1e-12 relative (max-norm and l2)."""
import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _check(u, ref, tol=TOL):
    assert np.isfinite(u).all()
    assert rel_max(u, ref) < tol, rel_max(u, ref)
    assert rel_l2(u, ref) < tol, rel_l2(u, ref)


def make_system(seed, n_fibers, n_shell, n_body_nodes, n_bodies, nodes=(8, 16, 24, 32, 48, 64)):
    rng = np.random.default_rng(seed)
    n_nodes = rng.choice(nodes, size=n_fibers) if n_fibers else np.zeros(0, dtype=int)
    lengths = rng.uniform(0.5, 2.0, n_fibers)
    pos = []
    for n, L in zip(n_nodes, lengths):
        x0 = rng.uniform(-3, 3, 3)
        nh = rng.normal(size=3)
        nh /= np.linalg.norm(nh)
        pos.append(x0 + np.linspace(0, L, n)[:, None] * nh)
    r_fib = np.concatenate(pos) if pos else np.zeros((0, 3))
    d = rng.normal(size=(n_shell, 3))
    d /= np.linalg.norm(d, axis=1)[:, None] if n_shell else 1
    r_shell = d * np.array([7.8, 4.16, 4.16])
    n_shell_ = -d
    centers = rng.uniform(-2, 2, (n_bodies, 3))
    per = n_body_nodes // max(n_bodies, 1)
    r_body, n_body = [], []
    for b in range(n_bodies):
        e = rng.normal(size=(per, 3))
        e /= np.linalg.norm(e, axis=1)[:, None]
        r_body.append(centers[b] + 0.4 * e)
        n_body.append(e)
    r_body = np.concatenate(r_body) if r_body else np.zeros((0, 3))
    n_body = np.concatenate(n_body) if n_body else np.zeros((0, 3))
    fib = dict(pos=r_fib, n_nodes=list(n_nodes), lengths=list(lengths), forces=rng.uniform(-1, 1, r_fib.shape))
    shell = dict(pos=r_shell, normals=n_shell_, density=rng.uniform(-1, 1, r_shell.shape))
    body = dict(pos=r_body, normals=n_body, density=rng.uniform(-1, 1, r_body.shape), centers=centers,
                forces=rng.uniform(-1, 1, (n_bodies, 3)), torques=rng.uniform(-1, 1, (n_bodies, 3)))
    return fib, shell, body


def load(fl, fib, shell, body):
    fl.set_fibers(fib["pos"], fib["n_nodes"], fib["lengths"])
    fl.set_periphery(shell["pos"], shell["normals"])
    fl.set_bodies(body["pos"], body["normals"], body["centers"])


def ft_of(body):
    return np.concatenate([body["forces"], body["torques"]], axis=1)


@pytest.mark.parametrize("eta", [1.0, 0.37])
def test_fiber_flow_with_self_subtraction(eta):
    fib, shell, body = make_system(1, 37, 300, 0, 0)
    r_all = np.concatenate([fib["pos"], shell["pos"]])
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v = fl.fiber_flow(r_all, fib["forces"], eta, subtract_self=True)
        v_ns = fl.fiber_flow(r_all, fib["forces"], eta, subtract_self=False)
    _check(v, orc.fiber_flow(r_all, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, True))
    _check(v_ns, orc.fiber_flow(r_all, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, False))


def test_periphery_and_body_flows():
    fib, shell, body = make_system(2, 10, 1500, 800, 2)
    rng = np.random.default_rng(9)
    r_trg = rng.uniform(-2, 2, (777, 3))
    eta = 1.3
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        vs = fl.periphery_flow(r_trg, shell["density"], eta)
        vb = fl.body_flow(r_trg, body["density"], ft_of(body), eta)
        # cached targets: second call with the same r_trg must give the same answer without re-upload
        vs2 = fl.periphery_flow(r_trg, shell["density"], eta)
    _check(vs, orc.periphery_flow(r_trg, shell["pos"], shell["normals"], shell["density"], eta))
    _check(vb, orc.body_flow(r_trg, body["pos"], body["normals"], body["density"], body["centers"], body["forces"],
                             body["torques"], eta))
    assert np.array_equal(vs, vs2)


@pytest.mark.parametrize("shape", [(16, 0, 0, 0), (16, 200, 0, 0), (20, 500, 400, 1), (0, 300, 400, 2),
                                   (12, 0, 600, 3), (0, 0, 0, 0), (60, 1000, 800, 2)])
def test_fused_matvec_flow(shape):
    n_fibers, n_shell, n_body_nodes, n_bodies = shape
    fib, shell, body = make_system(7 + sum(shape), n_fibers, n_shell, n_body_nodes, n_bodies)
    eta = 0.9
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        st = fl.stats()
        # GMRES: same geometry, new strengths
        rng = np.random.default_rng(0)
        fib2 = dict(fib, forces=rng.normal(size=fib["forces"].shape))
        v2 = fl.matvec(fib2["forces"], shell["density"], body["density"], ft_of(body), eta)
    ref = orc.matvec_flow(fib, shell, body, eta)
    assert v.shape == ref.shape
    if ref.size == 0:
        return
    if np.abs(ref).max() == 0.0:
        assert np.abs(v).max() == 0.0
        return
    _check(v, ref)
    _check(v2, orc.matvec_flow(fib2, shell, body, eta))
    assert st["launches"] > 0 and st["n_pairs"] > 0


def test_c1_plumbing_config_16x32():
    # BASELINE configs[0]: 16 fibers x 32 nodes, direct kernel, targets == sources
    fib, shell, body = make_system(4, 16, 0, 0, 0, nodes=(32,))
    eta = 1.0
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    _check(v, orc.matvec_flow(fib, shell, body, eta))


def test_geometry_update_between_timesteps():
    fib, shell, body = make_system(11, 8, 100, 0, 0)
    eta = 1.0
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v0 = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        fib2 = dict(fib, pos=fib["pos"] + 0.01)  # System::step moves the fibers (system.cpp:486-489)
        fl.set_fibers(fib2["pos"], fib2["n_nodes"], fib2["lengths"])
        v1 = fl.matvec(fib2["forces"], shell["density"], body["density"], ft_of(body), eta)
    _check(v0, orc.matvec_flow(fib, shell, body, eta))
    _check(v1, orc.matvec_flow(fib2, shell, body, eta))


def test_target_windows_tile_the_full_matvec():
    # one-rank-per-GPU sharding of apply_matvec: every rank evaluates a block of [fibers | shell | bodies] rows;
    # blocks cut through fibers, the shell and the body on purpose
    fib, shell, body = make_system(21, 25, 333, 450, 1)
    eta = 1.1
    ref = orc.matvec_flow(fib, shell, body, eta)
    n_all = ref.shape[0]
    cuts = [0, 37, 200, n_all // 2, n_all - 100, n_all]
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        full = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            fl.set_target_window(a, b)
            parts.append(fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta))
    _check(full, ref)
    tiled = np.concatenate(parts)
    assert tiled.shape == ref.shape
    _check(tiled, ref)


def test_matvec_device_pointers_with_torch():
    import torch
    fib, shell, body = make_system(23, 30, 400, 300, 2)
    eta = 0.8
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_ff, d_sd, d_bd = t(fib["forces"]), t(shell["density"]), t(body["density"])
    d_f, d_t = t(body["forces"]), t(body["torques"])
    n_all = fib["pos"].shape[0] + shell["pos"].shape[0] + body["pos"].shape[0]
    d_v = torch.empty((n_all, 3), dtype=torch.float64, device=dev)
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        st = torch.cuda.current_stream().cuda_stream
        fl.matvec_device(d_ff.data_ptr(), d_sd.data_ptr(), d_bd.data_ptr(), d_f.data_ptr(), d_t.data_ptr(), eta,
                         d_v.data_ptr(), st)
        torch.cuda.synchronize()
    _check(d_v.cpu().numpy(), orc.matvec_flow(fib, shell, body, eta))


@pytest.mark.parametrize("n_trg", [1, 6, 1000])
def test_velocity_at_targets(n_trg):
    # listener / streamline path: fiber flow without self subtraction + body flow + periphery flow at free targets
    # (system.cpp:355-359); 1-6 targets is what the Cash-Karp streamline integrator asks for per stage
    fib, shell, body = make_system(33, 20, 600, 300, 1)
    rng = np.random.default_rng(n_trg)
    r_trg = rng.uniform(-2.5, 2.5, (n_trg, 3))
    eta = 1.0
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        v = fl.velocity_at_targets(r_trg, fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    ref = orc.fiber_flow(r_trg, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, subtract_self=False)
    ref += orc.body_flow(r_trg, body["pos"], body["normals"], body["density"], body["centers"], body["forces"],
                         body["torques"], eta)
    ref += orc.periphery_flow(r_trg, shell["pos"], shell["normals"], shell["density"], eta)
    _check(v, ref)


def test_velocity_at_targets_with_point_and_background_sources():
    # + psc_.flow + bs_.flow (system.cpp:358-359; point_source.cpp:16-54, background_source.cpp:15-24)
    fib, shell, body = make_system(35, 10, 200, 0, 0)
    rng = np.random.default_rng(2)
    r_trg = rng.uniform(-2, 2, (500, 3))
    pts = rng.uniform(-1, 1, (3, 3))
    r_trg[0] = pts[0]                      # a target exactly on a point source: skipped, no NaN
    r_trg[1] = pts[1] + 1e-7               # inside the regularisation radius (eps = 1e-5)
    pf, ptq = rng.normal(size=(3, 3)), rng.normal(size=(3, 3))
    comp, scale, uni = [1, 0, 2], np.array([0.5, -0.25, 0.0]), np.array([0.1, 0.0, -0.3])
    eta = 0.8
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        fl.set_point_sources(pts, pf, ptq)
        fl.set_background(comp, scale, uni)
        v = fl.velocity_at_targets(r_trg, fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    ref = orc.fiber_flow(r_trg, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, subtract_self=False)
    ref += orc.periphery_flow(r_trg, shell["pos"], shell["normals"], shell["density"], eta)
    ref += orc.oseen_contract(pts, r_trg, pf, eta) + orc.rotlet(pts, r_trg, ptq, eta)
    ref += uni[None, :] + r_trg[:, comp] * scale[None, :]
    _check(v, ref)


def test_cpp_flow_engine():
    # tests/cpp/flow_test.cpp: C++ host code -> skelly_b200/flow.hpp -> C ABI, against host loops
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "flow_test")
    assert os.path.exists(exe), "tests/cpp/flow_test not built (run __graft_entry__.build())"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_small_calls_replayed_as_cuda_graphs():
    # launch-bound sizes (BASELINE C1, listener / streamline path) are captured on the third identical-shape call and
    # replayed as one CUDA graph afterwards; strengths, targets and geometry keep changing underneath
    fib, shell, body = make_system(41, 16, 120, 80, 1, nodes=(32,))
    rng = np.random.default_rng(4)
    eta = 1.1
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        for it in range(7):
            fib_i = dict(fib, forces=rng.normal(size=fib["forces"].shape))
            shell_i = dict(shell, density=rng.normal(size=shell["density"].shape))
            v = fl.matvec(fib_i["forces"], shell_i["density"], body["density"], ft_of(body), eta)
            _check(v, orc.matvec_flow(fib_i, shell_i, body, eta))
        launches_replayed = fl.stats()["launches"]
        assert launches_replayed > 0
        # streamline-like: 6 fresh targets every call
        for it in range(6):
            r_trg = rng.uniform(-2, 2, (6, 3))
            v = fl.velocity_at_targets(r_trg, fib["forces"], shell["density"], body["density"], ft_of(body), eta)
            ref = orc.fiber_flow(r_trg, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, False)
            ref += orc.body_flow(r_trg, body["pos"], body["normals"], body["density"], body["centers"],
                                 body["forces"], body["torques"], eta)
            ref += orc.periphery_flow(r_trg, shell["pos"], shell["normals"], shell["density"], eta)
            _check(v, ref)
        # a different target count in between (buffers may move), then back
        r_big = rng.uniform(-2, 2, (3000, 3))
        fl.velocity_at_targets(r_big, fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        for it in range(4):
            r_trg = rng.uniform(-2, 2, (6, 3))
            v = fl.velocity_at_targets(r_trg, fib["forces"], shell["density"], body["density"], ft_of(body), eta)
            ref = orc.fiber_flow(r_trg, fib["pos"], fib["n_nodes"], fib["lengths"], fib["forces"], eta, False)
            ref += orc.body_flow(r_trg, body["pos"], body["normals"], body["density"], body["centers"],
                                 body["forces"], body["torques"], eta)
            ref += orc.periphery_flow(r_trg, shell["pos"], shell["normals"], shell["density"], eta)
            _check(v, ref)
        # geometry update invalidates the matvec graph
        fib2 = dict(fib, pos=fib["pos"] + 0.02)
        fl.set_fibers(fib2["pos"], fib2["n_nodes"], fib2["lengths"])
        for it in range(4):
            v = fl.matvec(fib2["forces"], shell["density"], body["density"], ft_of(body), eta)
            _check(v, orc.matvec_flow(fib2, shell, body, eta))


def test_graphs_can_be_disabled(monkeypatch):
    monkeypatch.setenv("SKB_GRAPHS", "0")
    fib, shell, body = make_system(43, 8, 50, 0, 0, nodes=(16,))
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        for _ in range(4):
            v = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), 1.0)
    _check(v, orc.matvec_flow(fib, shell, body, 1.0))


@pytest.mark.parametrize("n_fibers,n_shell,n_body,n_bodies", [(130, 700, 360, 3), (130, 700, 0, 0), (90, 33, 100, 1)])
def test_matvec_cross_kernel_matches_the_two_call_form(n_fibers, n_shell, n_body, n_bodies, monkeypatch):
    """Fiber <-> periphery pairs in one geometry pass (cross_kernels.cuh: the fibers' Stokeslets on the periphery and the
    periphery's stresslets on the fibers share d, r^2, 1/r) against the oracle and against the two separate evaluator
    calls the reference issues (system.cpp:299, :304)."""
    monkeypatch.setenv("SKB_SYMMETRIC", "1")
    fib, shell, body = make_system(500 + n_fibers, n_fibers, n_shell, n_body, n_bodies, nodes=(16, 32, 48, 64, 96))
    eta = 0.6
    ref = orc.matvec_flow(fib, shell, body, eta)
    with skb.Flow(0) as fl:
        load(fl, fib, shell, body)
        fl.set_cross(0)
        v_two = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        fl.set_cross(1)
        v_cross = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        v_again = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
        fl.set_self_exclusion(True)
        v_excl = fl.matvec(fib["forces"], shell["density"], body["density"], ft_of(body), eta)
    _check(v_two, ref)
    _check(v_cross, ref)
    _check(v_excl, ref)
    assert np.array_equal(v_cross, v_again), "bitwise reproducible"
