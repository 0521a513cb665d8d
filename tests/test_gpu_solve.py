"""GPU test of the solve-shaped loop: K right-preconditioned GMRES iterations' worth of operator
applications -- P_inv_hydro::apply then A_fiber_hydro::apply (src/core/solver_hydro.cpp:23-48, system.cpp:248-324) --
with every vector resident on the device (skellysim_b200/solver_hydro.py over the C ABI), against the same K
iterations of the oracle's restatements.  Tolerance 1e-11 after K = 6 iterations (errors compound through the loop)."""
import numpy as np
import pytest

import oracle as orc
import skellysim_b200 as skb
from conftest import rel_max
from skellysim_b200.solver_hydro import HydroOperator, iterate
from test_gpu_fiberops import load_ops, make_ops
from test_gpu_flow import load, make_system

pytestmark = pytest.mark.gpu


def test_k_device_iterations_equal_k_oracle_iterations():
    import torch
    fib, shell, body = make_system(41, 70, 260, 150, 1, nodes=(16, 24, 32, 48))
    ops = make_ops(fib, 3)
    n_nodes = ops["n_nodes"]
    rng = np.random.default_rng(8)
    # well-conditioned fiber blocks (identity + small perturbation): LU solve and explicit inverse agree to rounding
    ops["A"] = [np.eye(4 * n) + 0.1 * A for n, A in zip(n_nodes, ops["A"])]
    nf, ns, nb = fib["pos"].shape[0], shell["pos"].shape[0], body["pos"].shape[0]
    M = rng.normal(size=(3 * ns, 3 * ns)) / np.sqrt(3 * ns)
    Minv = np.eye(3 * ns) + 0.1 * rng.normal(size=(3 * ns, 3 * ns)) / np.sqrt(3 * ns)
    eta, K = 1.2, 6
    x_f0, x_s0 = rng.normal(size=4 * nf), rng.normal(size=(ns, 3))
    link = rng.normal(size=(len(n_nodes), 7))

    # ---- oracle loop
    x_f, x_s = x_f0.copy(), x_s0.copy()
    for _ in range(K):
        y_f = orc.fiber_apply_preconditioner(ops["A"], x_f, n_nodes)          # fcfd.cpp:331-339
        y_s = (Minv @ x_s.reshape(-1)).reshape(ns, 3)                         # periphery.cpp:21-30
        res, v_all = orc.apply_matvec_fibers(fib, dict(shell, density=y_s), body, ops, y_f, eta, link)
        r_s = (M @ y_s.reshape(-1)).reshape(ns, 3) + v_all[nf:nf + ns]        # periphery.cpp:38-47
        scale = max(np.abs(res).max(), np.abs(r_s).max())
        x_f, x_s = res / scale, r_s / scale

    # ---- device loop
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    with skb.Flow(0) as fl, skb.Dense(device_ids=[0]) as dn:
        load(fl, fib, shell, body)
        load_ops(fl, ops)
        fl.set_fiber_preconditioner([np.linalg.inv(A) for A in ops["A"]])
        dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, M)
        dn.set_matrix(skb.DENSE_M_INV, Minv)
        op = HydroOperator(fl, dn, nf, ns, nb, eta, dev)
        d_f, d_s = iterate(op, t(x_f0), t(x_s0), t(body["density"]), t(body["forces"]), t(body["torques"]), t(link), K)
        torch.cuda.synchronize()
        got_f, got_s = d_f.cpu().numpy(), d_s.cpu().numpy()
        assert op.launches > 0
    assert rel_max(got_f, x_f) < 1e-11, rel_max(got_f, x_f)
    assert rel_max(got_s, x_s) < 1e-11, rel_max(got_s, x_s)
