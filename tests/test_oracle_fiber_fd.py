"""CPU tests that pin oracle/fiber_fd.py -- the restatement of FiberFiniteDifference::matrices_
(fiber_finite_difference.cpp:519-558, utils.cpp:12-102) and of the per-fiber operators -- as far as this container
allows (the reference's fiber code cannot be compiled here and stores no golden numbers): literal statements of the
reference, and the defining properties of the formulas it cites."""
import numpy as np
import pytest

from oracle import fiber_fd as ffd
import oracle as orc


@pytest.mark.parametrize("n", ffd.ALLOWED_N_NODES)
def test_weights_and_downsampling_layout_follow_the_reference_statements(n):
    m = ffd.compute_matrices(n)
    # weights_0: 2 everywhere, 1 at the ends, / (n - 1)      fiber_finite_difference.cpp:545-548
    w = m["weights_0"]
    assert w.shape == (n,) and np.isclose(w.sum(), 2.0)
    assert np.allclose(w[1:-1], 2.0 / (n - 1)) and w[0] == w[-1] == 1.0 / (n - 1)
    assert np.allclose(orc.trapezoid_weights(n, 1.7), 0.5 * 1.7 * w)
    # P_downsample_bc: three P_X blocks of (n-4) x n and one P_T of (n-2) x n on the block diagonal   :551-555
    P = m["P_downsample_bc"]
    assert P.shape == (4 * n - 14, 4 * n)
    mask = np.zeros_like(P, dtype=bool)
    for k in range(3):
        mask[k * (n - 4):(k + 1) * (n - 4), k * n:(k + 1) * n] = True
        assert np.array_equal(P[k * (n - 4):(k + 1) * (n - 4), k * n:(k + 1) * n], m["P_X"])
    mask[3 * (n - 4):, 3 * n:] = True
    assert np.array_equal(P[3 * (n - 4):, 3 * n:], m["P_T"])
    assert not P[~mask].any()                      # 75 % structural zeros
    assert m["P_X"].shape == (n - 4, n) and m["P_T"].shape == (n - 2, n)
    # the grids                                                                                    :524-531
    assert np.allclose(m["alpha"], np.linspace(-1, 1, n))
    assert np.allclose(m["alpha_roots"], (2 * np.arange(n - 4) + 1) / (n - 4) - 1)
    assert np.allclose(m["alpha_tension"], (2 * np.arange(n - 2) + 1) / (n - 2) - 1)


@pytest.mark.parametrize("n", [8, 16, 32, 64])
@pytest.mark.parametrize("M", [1, 2, 3, 4])
def test_finite_diff_is_exact_on_polynomials_up_to_the_stencil_degree(n, M):
    # Fornberg's weights on an n_s-point stencil differentiate polynomials of degree < n_s exactly (utils.cpp:38-47)
    s = np.linspace(-1, 1, n)
    n_s = 4 + M                                     # compute_matrices_finitediff(4): :537-540
    D = ffd.finite_diff(s, M, n_s)
    assert D.shape == (n, n)
    assert (np.count_nonzero(D, axis=1) <= n_s).all()              # banded: a sliding window of n_s points
    rng = np.random.default_rng(n + M)
    for deg in range(0, n_s):
        p = np.polynomial.Polynomial(rng.normal(size=deg + 1))
        got, want = D @ p(s), p.deriv(M)(s)
        assert np.abs(got - want).max() <= 1e-7 * max(1.0, np.abs(want).max()) * (n / 8) ** M
    # and not beyond (the test would be vacuous if D were, say, spectrally accurate)
    p = np.polynomial.Polynomial([0] * (n_s + 1) + [1.0])
    assert np.abs(D @ p(s) - p.deriv(M)(s)).max() > 1e-6
    assert np.array_equal(ffd.compute_matrices(n)[f"D_{M}_0"], D.T)  # stored pre-transposed (:537)


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
def test_barycentric_matrix_interpolates(n):
    x = np.linspace(-1, 1, n)
    y = (2 * np.arange(n - 4) + 1) / (n - 4) - 1
    P = ffd.barycentric_matrix(x, y)
    assert np.allclose(P.sum(axis=1), 1.0, atol=1e-13)              # reproduces constants (a rational interpolant)
    # a point that coincides with a node picks that node's value (utils.cpp:29-32)
    P2 = ffd.barycentric_matrix(x, x[[0, 3, n - 1]])
    assert np.allclose(P2 @ np.cos(x), np.cos(x[[0, 3, n - 1]]), atol=1e-12)
    # and it converges on a smooth function
    f = lambda t: np.sin(2.0 * t) + 0.3 * t ** 2
    assert np.abs(P @ f(x) - f(y)).max() < 5.0 / n ** 2


@pytest.mark.parametrize("n", [16, 32])
def test_operators_of_a_straight_fiber(n):
    L = 1.3
    nh = np.array([1.0, 2.0, -0.5])
    nh /= np.linalg.norm(nh)
    x = np.array([0.2, -0.1, 0.4]) + np.linspace(0, L, n)[:, None] * nh
    fo = ffd.FiberOperators(x, L)
    assert np.allclose(fo.xs, nh, atol=1e-11)                       # unit tangent (update_derivatives, :62-72)
    assert np.abs(fo.xss).max() < 1e-8 and np.abs(fo.xsss).max() < 1e-6
    assert fo.A.shape == (4 * n, 4 * n) and fo.force_operator.shape == (3 * n, 4 * n)
    # force_operator_ [x; y; z; T]: bending force -E x_ssss vanishes on the straight shape, and a uniform tension T
    # gives (T x_s)_s = T x_ss = 0 in the interior (update_force_operator, :317-335)
    X = np.concatenate([x[:, 0], x[:, 1], x[:, 2], np.zeros(n)])
    assert np.abs(fo.force_operator @ X).max() < 1e-4
    Tn = np.concatenate([np.zeros(3 * n), np.ones(n)])
    f = (fo.force_operator @ Tn).reshape(3, n)
    assert np.abs(f[:, 3:-3]).max() < 1e-9
    # free ends: the last 14 rows are the boundary conditions (Force / Torque at both ends, :386-433, :466-505)
    B = fo.A[4 * n - 14:]
    assert np.count_nonzero(B[4:7].sum(axis=0)) > 0 and B[3, 3 * n] == -1 and B[10, 4 * n - 1] == 1.0
    # clamped minus end / pinned plus end switch the other branches on
    fc = ffd.FiberOperators(x, L, minus_clamped=True, plus_pinned=True)
    assert fc.A[4 * n - 14, 0] == fc.beta / fc.dt and fc.plus_bc_velocity == 1 and fo.plus_bc_velocity == 0


def test_real_operators_through_the_matvec_restatements():
    # oracle.fiber_matvec (loop form, ffd.cpp:276-312) against the independently assembled velocity operator, now on
    # reference-shaped operators instead of random stand-ins
    rng = np.random.default_rng(3)
    n, L = 24, 0.9
    t = np.linspace(0, L, n)
    x = np.stack([t, 0.05 * np.sin(3 * t), 0.03 * np.cos(2 * t)], axis=1)
    fo = ffd.FiberOperators(x, L, plus_pinned=True)
    m = fo.mats
    xv, v, vb = rng.normal(size=4 * n), rng.normal(size=(n, 3)), rng.normal(size=7)
    got = orc.fiber_matvec(fo.A, m["D_1_0"], m["P_downsample_bc"], fo.xs, fo.length_prev, fo.plus_bc_velocity, xv, v, vb)
    V = orc.fiber_velocity_operator(m["D_1_0"], m["P_downsample_bc"], fo.xs, fo.length_prev, fo.plus_bc_velocity)
    y_bc = np.zeros(4 * n)
    y_bc[4 * n - 14:4 * n - 7] = vb
    want = fo.A @ xv + V @ v.reshape(-1) + y_bc
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    ops = ffd.suspension_operators(np.concatenate([x, x + 1.0]), [n, n], [L, L])
    assert len(ops["A"]) == 2 and ops["xs"].shape == (2 * n, 3) and set(ops["D_1_0"]) == {n}
