"""ctypes front-end of oracle/_ref/liboracle.so plus numpy restatements of the flow() layer.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Array convention follows the reference:
a ``3 x n`` Eigen column-major matrix is passed here as a C-contiguous ``(n, 3)`` float64 array
(identical bytes); a ``9 x n`` stresslet strength is ``(n, 9)`` with ``f[i, 3*a + b]``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "liboracle.so")
_REFGPU_PATH = os.path.join(_HERE, "_ref", "libskelly_ref_kernels_cu.so")

_dp = C.POINTER(C.c_double)
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c (and, when /root/reference is present, the reference's kernels.cu) via oracle/Makefile."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        for name in ("oracle_stokeslet_direct", "oracle_stresslet_direct", "oracle_stokeslet_direct_ld",
                     "oracle_stresslet_direct_ld"):
            fn = getattr(L, name)
            fn.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int]
            fn.restype = None
        for name in ("oracle_stokeslet_direct_cpu", "oracle_stresslet_direct_cpu"):
            fn = getattr(L, name)
            fn.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int]
            fn.restype = C.c_int
        L.oracle_oseen_contract.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_double,
                                            C.c_double, _dp]
        L.oracle_oseen_contract.restype = None
        L.oracle_rotlet_add.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_double, _dp]
        L.oracle_rotlet_add.restype = None
        L.oracle_form_double_layer.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp]
        L.oracle_form_double_layer.restype = None
        L.oracle_simd_level.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _a(x, cols):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.size == 0:
        return x.reshape(0, cols)
    assert x.ndim == 2 and x.shape[1] == cols, (x.shape, cols)
    return x


def _p(x):
    return x.ctypes.data_as(_dp)


def _direct(name, r_src, f_src, r_trg, fdim):
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, fdim), _a(r_trg, 3)
    assert r_src.shape[0] == f_src.shape[0]
    u = np.empty((r_trg.shape[0], 3))
    getattr(lib(), name)(_p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0])
    return u


def stokeslet_direct(r_src, f_src, r_trg):
    """kernels.cu:57-77,79-123 semantics: includes 1/(8 pi), excludes 1/eta, r = 0 pairs skipped."""
    return _direct("oracle_stokeslet_direct", r_src, f_src, r_trg, 3)


def stresslet_direct(r_src, f_src, r_trg):
    """kernels.cu:24-55 semantics (9 strengths per source)."""
    return _direct("oracle_stresslet_direct", r_src, f_src, r_trg, 9)


def stokeslet_direct_ld(r_src, f_src, r_trg):
    return _direct("oracle_stokeslet_direct_ld", r_src, f_src, r_trg, 3)


def stresslet_direct_ld(r_src, f_src, r_trg):
    return _direct("oracle_stresslet_direct_ld", r_src, f_src, r_trg, 9)


def _direct_cpu(name, r_src, f_src, r_trg, fdim, eta, n_threads, simd):
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, fdim), _a(r_trg, 3)
    u = np.empty((r_trg.shape[0], 3))
    rc = getattr(lib(), name)(_p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0], float(eta),
                              int(n_threads), int(simd))
    if rc < 0:
        raise RuntimeError("requested SIMD level not supported by this CPU")
    return u


def stokeslet_direct_cpu(r_src, f_src, r_trg, eta, n_threads=0, simd=-1):
    """kernels::stokeslet_direct_cpu (kernels.cpp:54-67): OpenMP target chunks, result / eta."""
    return _direct_cpu("oracle_stokeslet_direct_cpu", r_src, f_src, r_trg, 3, eta, n_threads, simd)


def stresslet_direct_cpu(r_src, f_src, r_trg, eta, n_threads=0, simd=-1):
    """kernels::stresslet_direct_cpu (kernels.cpp:69-83)."""
    return _direct_cpu("oracle_stresslet_direct_cpu", r_src, f_src, r_trg, 9, eta, n_threads, simd)


def stokeslet_direct_excluding(r_src, f_src, r_trg, ids):
    """Stokeslet sum where the targets START with the sources and pairs (target i < n_src, source j) with
    ids[i] == ids[j] are left out: the fused form of FiberContainerFiniteDifference::flow's "all pairs, then subtract the
    fiber's own block" (fiber_container_finite_difference.cpp:203-210) for fibers whose nodes are farther apart than the
    regularisation threshold 1e-5 (kernels.cpp:176-184).  Built as (all pairs) - (same-id pairs), each by the scalar
    restatement of kernels.cu:57-77."""
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, 3), _a(r_trg, 3)
    ids = np.asarray(ids).reshape(-1)
    u = stokeslet_direct(r_src, f_src, r_trg)
    for g in np.unique(ids):
        m = np.nonzero(ids == g)[0]
        u[m] -= stokeslet_direct(r_src[m], f_src[m], r_src[m])
    return u


def simd_level() -> int:
    return lib().oracle_simd_level()


def max_threads() -> int:
    return lib().oracle_max_threads()


def form_double_layer(normals, density, eta):
    """periphery.cpp:68-71 / body_container.cpp:296-302: f_dl[i, 3a+b] = 2 eta n[i,a] rho[i,b]."""
    normals, density = _a(normals, 3), _a(density, 3)
    f = np.empty((normals.shape[0], 9))
    lib().oracle_form_double_layer(_p(normals), _p(density), normals.shape[0], float(eta), _p(f))
    return f


def oseen_contract(r_src, r_trg, density, eta, reg=5e-3, eps=1e-5):
    """kernels::oseen_tensor_direct(r_src, r_trg, eta) @ density without forming the matrix (kernels.cpp:146-195)."""
    r_src, r_trg, density = _a(r_src, 3), _a(r_trg, 3), _a(density, 3)
    out = np.zeros((r_trg.shape[0], 3))
    lib().oracle_oseen_contract(_p(r_src), r_src.shape[0], _p(r_trg), r_trg.shape[0], _p(density), float(eta),
                                float(reg), float(eps), 1.0, _p(out))
    return out


def rotlet(r_src, r_trg, density, eta, reg=5e-3, eps=1e-5):
    """kernels::rotlet (kernels.cpp:206-242)."""
    r_src, r_trg, density = _a(r_src, 3), _a(r_trg, 3), _a(density, 3)
    out = np.zeros((r_trg.shape[0], 3))
    lib().oracle_rotlet_add(_p(r_src), r_src.shape[0], _p(r_trg), r_trg.shape[0], _p(density), float(eta),
                            float(reg), float(eps), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# numpy restatements (independent of the C code; small sizes only) -- used to cross-check oracle.c
# ----------------------------------------------------------------------------------------------

def stokeslet_direct_numpy(r_src, f_src, r_trg):
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, 3), _a(r_trg, 3)
    d = r_trg[:, None, :] - r_src[None, :, :]
    r2 = np.einsum("tsk,tsk->ts", d, d)
    with np.errstate(divide="ignore"):
        rinv = np.where(r2 == 0.0, 0.0, 1.0 / np.sqrt(r2))
    inner = np.einsum("tsk,sk->ts", d, f_src) * rinv * rinv
    u = np.einsum("ts,sk->tk", rinv, f_src) + np.einsum("ts,tsk->tk", rinv * inner, d)
    return u / (8.0 * np.pi)


def stresslet_direct_numpy(r_src, f_src, r_trg):
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, 9), _a(r_trg, 3)
    d = r_trg[:, None, :] - r_src[None, :, :]
    r2 = np.einsum("tsk,tsk->ts", d, d)
    with np.errstate(divide="ignore"):
        rinv = np.where(r2 == 0.0, 0.0, 1.0 / np.sqrt(r2))
    S = f_src.reshape(-1, 3, 3)
    coeff = np.einsum("tsa,sab,tsb->ts", d, S, d) * (-3.0) * rinv**5
    return np.einsum("ts,tsk->tk", coeff, d) / (8.0 * np.pi)


# ----------------------------------------------------------------------------------------------
# flow() layer restatements (reference semantics, host arithmetic) on top of the C pair kernels
# ----------------------------------------------------------------------------------------------

def trapezoid_weights(n_nodes: int, length: float):
    """0.5 * L * weights_0 (fiber_container_finite_difference.cpp:186, fiber_finite_difference.cpp:545-548)."""
    w = np.full(n_nodes, 2.0)
    w[0] = w[-1] = 1.0
    w /= (n_nodes - 1)          # weights_0 /= (n_nodes - 1)        fiber_finite_difference.cpp:548
    return (0.5 * length) * w   # 0.5 * fib.length_ * weights_0     fiber_container_finite_difference.cpp:186


def fiber_flow(r_trg, fiber_pos, fiber_n_nodes, fiber_lengths, fib_forces, eta, subtract_self=True):
    """FiberContainerFiniteDifference::flow (fiber_container_finite_difference.cpp:172-214).

    fiber_pos (N_f,3) concatenated node positions; fiber_n_nodes / fiber_lengths per fiber;
    fib_forces (N_f,3).  When subtract_self, the first N_f targets ARE the fiber nodes (apply_matvec order).
    """
    fiber_pos, fib_forces, r_trg = _a(fiber_pos, 3), _a(fib_forces, 3), _a(r_trg, 3)
    if len(fiber_n_nodes) == 0:
        return np.zeros((r_trg.shape[0], 3))
    w = np.concatenate([trapezoid_weights(n, L) for n, L in zip(fiber_n_nodes, fiber_lengths)])
    wf = fib_forces * w[:, None]
    vel = stokeslet_direct(fiber_pos, wf, r_trg) / eta
    if subtract_self:
        off = 0
        for n in fiber_n_nodes:
            x = fiber_pos[off:off + n]
            vel[off:off + n] -= oseen_contract(x, x, wf[off:off + n], eta)
            off += n
    return vel


def periphery_flow(r_trg, node_pos, node_normal, density, eta):
    """Periphery::flow (periphery.cpp:55-79)."""
    r_trg = _a(r_trg, 3)
    if _a(node_pos, 3).shape[0] == 0:
        return np.zeros((r_trg.shape[0], 3))
    f_dl = form_double_layer(node_normal, density, eta)
    return stresslet_direct(node_pos, f_dl, r_trg) / eta


def body_flow(r_trg, node_pos, node_normal, densities, centers, forces, torques, eta):
    """BodyContainer::flow_spherical / flow_ellipsoidal (body_container.cpp:269-339, 341-411): stresslet from the
    body nodes + Stokeslet from the centres (net forces) + rotlet from the centres (net torques)."""
    r_trg = _a(r_trg, 3)
    if _a(node_pos, 3).shape[0] == 0:
        return np.zeros((r_trg.shape[0], 3))
    f_dl = form_double_layer(node_normal, densities, eta)
    v = stresslet_direct(node_pos, f_dl, r_trg) / eta
    v += stokeslet_direct(centers, forces, r_trg) / eta
    v += rotlet(centers, r_trg, torques, eta)
    return v


def periphery_dense_apply(A, x, v_add=None):
    """Periphery::matvec `stresslet_plus_complementary_ * x_shell + v_local` (periphery.cpp:38-47) and
    Periphery::apply_preconditioner `M_inv_ * x_shell` (periphery.cpp:21-30); A row-major as the precompute .npz
    stores it (precompute.py:113-148)."""
    y = np.asarray(A, dtype=np.float64) @ np.asarray(x, dtype=np.float64).reshape(-1)
    if v_add is not None:
        y = y + np.asarray(v_add, dtype=np.float64).reshape(-1)
    return y


def matvec_flow(fib, shell, body, eta):
    """Hydrodynamic part of System::apply_matvec (system.cpp:284-316): v_all over targets
    [fibers | shell | bodies]; shell sources act on fiber and body targets only (system.cpp:301-315).

    fib   = dict(pos, n_nodes, lengths, forces)
    shell = dict(pos, normals, density)
    body  = dict(pos, normals, density, centers, forces, torques)
    """
    r_f, r_s, r_b = _a(fib["pos"], 3), _a(shell["pos"], 3), _a(body["pos"], 3)
    nf, ns, nb = r_f.shape[0], r_s.shape[0], r_b.shape[0]
    r_all = np.concatenate([r_f, r_s, r_b])
    v_all = fiber_flow(r_all, r_f, fib["n_nodes"], fib["lengths"], fib["forces"], eta, subtract_self=True)
    r_fb = np.concatenate([r_f, r_b])
    v_s = periphery_flow(r_fb, r_s, shell["normals"], shell["density"], eta)
    v_all[:nf] += v_s[:nf]
    v_all[nf + ns:] += v_s[nf:]
    v_all += body_flow(r_all, r_b, body["normals"], body["density"], body["centers"], body["forces"],
                       body["torques"], eta)
    return v_all


# ----------------------------------------------------------------------------------------------
# per-fiber dense operators of the matvec.  PARITY UNPINNED: the reference's fiber code needs
# Eigen/toml/spdlog/MPI and cannot be built here, and its Python package has no counterpart, so these numpy
# restatements are checked only against each other (loop form vs. assembled dense operator), not against
# reference output.
# ----------------------------------------------------------------------------------------------

def apply_fiber_force(force_ops, x_fibers, n_nodes):
    """FiberContainerFiniteDifference::apply_fiber_force (fiber_container_finite_difference.cpp:272-287).

    force_ops: list of (3n, 4n) arrays (FiberFiniteDifference::force_operator_); x_fibers: concatenated per-fiber
    [x; y; z; T] of 4n.  Returns fw as (N_f, 3) (the reference's 3 x N_f, column-major)."""
    x_fibers = np.asarray(x_fibers, dtype=np.float64).reshape(-1)
    fw = np.zeros((int(np.sum(n_nodes)), 3))
    off = 0
    for F, n in zip(force_ops, n_nodes):
        ff = np.asarray(F) @ x_fibers[4 * off:4 * off + 4 * n]     # :278
        for k in range(3):
            fw[off:off + n, k] = ff[k * n:(k + 1) * n]             # :279-281
        off += n
    return fw


def fiber_matvec(A, D_1_0, P_downsample_bc, xs, length_prev, plus_bc_velocity, x, v, v_boundary=None):
    """FiberFiniteDifference::matvec (fiber_finite_difference.cpp:276-312) for one fiber.

    A (4n,4n); D_1_0 (n,n); P_downsample_bc (4n-14,4n); xs (n,3) tangents; x (4n); v (n,3); v_boundary (7,) or
    None."""
    n = xs.shape[0]
    bc = 4 * n - 14                                                 # :279
    D_1 = np.asarray(D_1_0) * (2.0 / length_prev)                   # :280
    xsDs = (D_1 * xs[:, 0][:, None]).T                              # :281  (D_1.colwise() * xs_x)^T
    ysDs = (D_1 * xs[:, 1][:, None]).T
    zsDs = (D_1 * xs[:, 2][:, None]).T
    vT = np.empty(4 * n)
    vT[0:n], vT[n:2 * n], vT[2 * n:3 * n] = v[:, 0], v[:, 1], v[:, 2]   # :290-292
    vT[3 * n:] = xsDs @ v[:, 0] + ysDs @ v[:, 1] + zsDs @ v[:, 2]   # :293
    vT_in = np.zeros(4 * n)
    vT_in[:bc] = np.asarray(P_downsample_bc) @ vT                   # :296
    xs_vT = np.zeros(4 * n)
    xs_vT[bc + 3] = v[0] @ xs[0]                                    # :301
    y_BC = np.zeros(4 * n)
    if v_boundary is not None and np.size(v_boundary) > 0:
        y_BC[bc:bc + 7] = v_boundary                                # :305-306
    if plus_bc_velocity:
        xs_vT[bc + 10] = v[n - 1] @ xs[n - 1]                       # :308-309
    return np.asarray(A) @ np.asarray(x) - vT_in + xs_vT + y_BC     # :311


def fiber_container_matvec(ops, x_fibers, v_fibers, v_fib_boundary=None):
    """FiberContainerFiniteDifference::matvec (fiber_container_finite_difference.cpp:216-232).

    ops = dict(A=[...], D_1_0={n: ..}, P={n: ..}, xs (N_f,3), length_prev, plus, n_nodes);
    v_fib_boundary (n_fibers, 7) (the reference's 7 x n_fibers, column-major) or None."""
    x_fibers = np.asarray(x_fibers, dtype=np.float64).reshape(-1)
    res = np.zeros_like(x_fibers)
    off = 0
    for i, n in enumerate(ops["n_nodes"]):
        vb = None if v_fib_boundary is None else np.asarray(v_fib_boundary)[i]
        res[4 * off:4 * off + 4 * n] = fiber_matvec(
            ops["A"][i], ops["D_1_0"][n], ops["P"][n], ops["xs"][off:off + n], ops["length_prev"][i],
            ops["plus"][i], x_fibers[4 * off:4 * off + 4 * n], np.asarray(v_fibers)[off:off + n], vb)
        off += n
    return res


def fiber_apply_preconditioner(A_list, x_fibers, n_nodes):
    """FiberContainerFiniteDifference::apply_preconditioner (fiber_container_finite_difference.cpp:331-339):
    y_f = A_LU_.solve(x_f) per fiber, A_LU_ = PartialPivLU(A_) (fiber_finite_difference.cpp:340) -- LAPACK's
    partial-pivoting solve here."""
    x_fibers = np.asarray(x_fibers, dtype=np.float64).reshape(-1)
    y = np.empty_like(x_fibers)
    off = 0
    for A, n in zip(A_list, n_nodes):
        y[4 * off:4 * off + 4 * n] = np.linalg.solve(np.asarray(A), x_fibers[4 * off:4 * off + 4 * n])
        off += n
    return y


def fiber_velocity_operator(D_1_0, P_downsample_bc, xs, length_prev, plus_bc_velocity):
    """The (4n x 3n) matrix V with fiber_matvec(...) == A x + V vec(v) + y_BC, vec(v) = AoS [v_0x v_0y v_0z v_1x ...].
    An independent assembly of ffd.cpp:280-309 used to cross-check fiber_matvec."""
    n = xs.shape[0]
    bc = 4 * n - 14
    D_1 = np.asarray(D_1_0) * (2.0 / length_prev)
    T = np.zeros((4 * n, 3 * n))                    # vT = T vec(v)
    for i in range(n):
        for k in range(3):
            T[k * n + i, 3 * i + k] = 1.0
            T[3 * n:, 3 * i + k] = D_1[i, :] * xs[i, k]
    V = np.zeros((4 * n, 3 * n))
    V[:bc] = -np.asarray(P_downsample_bc) @ T
    V[bc + 3, 0:3] += xs[0]
    if plus_bc_velocity:
        V[bc + 10, 3 * (n - 1):3 * n] += xs[n - 1]
    return V


def apply_matvec_fibers(fib, shell, body, ops, x_fibers, eta, fiber_link_conditions=None):
    """System::apply_matvec (system.cpp:298-318) up to res_fibers: fw = apply_fiber_force(x_fibers), v_all =
    matvec_flow, res_fibers = fc.matvec(x_fibers, v_fibers, fiber_link_conditions).  Returns (res_fibers, v_all)."""
    fw = apply_fiber_force(ops["force"], x_fibers, ops["n_nodes"])
    fib = dict(fib, forces=fw)
    v_all = matvec_flow(fib, shell, body, eta)
    nf = int(np.sum(ops["n_nodes"]))
    res = fiber_container_matvec(ops, x_fibers, v_all[:nf], fiber_link_conditions)
    return res, v_all


# ----------------------------------------------------------------------------------------------
# the reference's own CUDA direct kernels (GPU box only): oracle/_ref/libskelly_ref_kernels_cu.so
# ----------------------------------------------------------------------------------------------
_refgpu = None


def refgpu_available() -> bool:
    return os.path.exists(_REFGPU_PATH)


def _refgpu_lib():
    global _refgpu
    if _refgpu is None:
        L = C.CDLL(_REFGPU_PATH)
        for mangled in ("_ZN7kernels25stokeslet_direct_gpu_implEPKdS1_iS1_Pdi",
                        "_ZN7kernels25stresslet_direct_gpu_implEPKdS1_iS1_Pdi"):
            fn = getattr(L, mangled)
            fn.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int]
            fn.restype = None
        _refgpu = L
    return _refgpu


def ref_stokeslet_direct_gpu_impl(r_src, f_src, r_trg):
    """kernels::stokeslet_direct_gpu_impl of the UNMODIFIED reference (src/core/kernels.cu:180-183)."""
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, 3), _a(r_trg, 3)
    u = np.zeros((r_trg.shape[0], 3))
    _refgpu_lib()._ZN7kernels25stokeslet_direct_gpu_implEPKdS1_iS1_Pdi(
        _p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0])
    return u


def ref_stresslet_direct_gpu_impl(r_src, f_src, r_trg):
    """kernels::stresslet_direct_gpu_impl of the UNMODIFIED reference (src/core/kernels.cu:185-188)."""
    r_src, f_src, r_trg = _a(r_src, 3), _a(f_src, 9), _a(r_trg, 3)
    u = np.zeros((r_trg.shape[0], 3))
    _refgpu_lib()._ZN7kernels25stresslet_direct_gpu_implEPKdS1_iS1_Pdi(
        _p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0])
    return u
