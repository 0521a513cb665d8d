"""ctypes binding of include/skelly_b200.h.

Array convention = the reference's: an Eigen column-major ``3 x n`` matrix is a C-contiguous ``(n, 3)``
float64 numpy array here (same bytes); stresslet strengths are ``(n, 9)``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

KERNEL_STOKESLET = 0
KERNEL_STRESSLET = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("SKB_LIBRARY") or os.path.join(_HERE, "lib", "libskelly_b200.so")  # override: tuning A/B only
_dp = C.POINTER(C.c_double)
_lib = None
BOUND_FUNCTIONS = ()  # names of the C-ABI functions this binding declares signatures for (filled by library())


class SkbError(RuntimeError):
    """Raised for every non-zero status of the C ABI (mirrors the std::runtime_error of the C++ wrapper)."""


class EvalStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("n_pairs", C.c_int64),
                ("launches", C.c_int32), ("targets_per_thread", C.c_int32), ("source_splits", C.c_int32),
                ("grid_ctas", C.c_int32)]


class FlowStats(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("total_ms", C.c_double), ("n_pairs", C.c_int64), ("launches", C.c_int32)]


class DenseStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("bytes", C.c_int64)]


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile libskelly_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "clean"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run(["make", "-C", src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libskelly_b200.so failed:\n" + r.stdout[-4000:])
    return _LIB


def library() -> C.CDLL:
    """Load the CUDA library.  Missing library == hard error: there is no other implementation to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise SkbError(f"{_LIB} is missing: build it with skellysim_b200.build_library() "
                       "(nvcc, sm_100a). There is no CPU or PyTorch fallback.")
    L = C.CDLL(_LIB)
    ctxp = C.c_void_p
    sig = {
        "skb_version": ([], C.c_char_p),
        "skb_last_error_string": ([], C.c_char_p),
        "skb_device_count": ([C.POINTER(C.c_int)], C.c_int),
        "skb_stokeslet_direct": ([_dp, _dp, C.c_int, _dp, _dp, C.c_int], C.c_int),
        "skb_stresslet_direct": ([_dp, _dp, C.c_int, _dp, _dp, C.c_int], C.c_int),
        "skb_ctx_create": ([C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_ctx_create_on": ([C.POINTER(C.c_int), C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_ctx_destroy": ([ctxp], C.c_int),
        "skb_ctx_n_gpus": ([ctxp, C.POINTER(C.c_int)], C.c_int),
        "skb_set_targets": ([ctxp, _dp, C.c_int64], C.c_int),
        "skb_set_sources": ([ctxp, C.c_int, _dp, C.c_int64], C.c_int),
        "skb_eval": ([ctxp, C.c_int, _dp, _dp, C.c_int], C.c_int),
        "skb_eval_fused": ([ctxp, _dp, _dp, _dp], C.c_int),
        "skb_set_source_normals": ([ctxp, _dp, C.c_int64], C.c_int),
        "skb_eval_double_layer": ([ctxp, _dp, C.c_double, _dp, C.c_int], C.c_int),
        "skb_set_source_exclusion_ids": ([ctxp, C.POINTER(C.c_int32), C.c_int64], C.c_int),
        "skb_set_targets_device": ([ctxp, C.c_void_p, C.c_int64, C.c_void_p], C.c_int),
        "skb_set_sources_device": ([ctxp, C.c_int, C.c_void_p, C.c_int64, C.c_void_p], C.c_int),
        "skb_eval_device": ([ctxp, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], C.c_int),
        "skb_sync": ([ctxp], C.c_int),
        "skb_last_eval_stats": ([ctxp, C.POINTER(EvalStats)], C.c_int),
        "skb_launch_count": ([], C.c_int64),
        "skb_ctx_set_tuning": ([ctxp, C.c_int, C.c_int], C.c_int),
        "skb_measure_fp64_peak": ([ctxp, C.POINTER(C.c_double)], C.c_int),
        "skb_ctx_set_symmetric": ([ctxp, C.c_int], C.c_int),
        "skb_plan_query": ([C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
                           + [C.POINTER(C.c_int)] * 4, C.c_int),
        "skb_sym_plan_query": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)], C.c_int),
        "skb_sym_groups_per_block": ([], C.c_int),
        "skb_ctx_set_sym_partition": ([ctxp, C.c_int, C.c_int], C.c_int),
        "skb_ctx_last_sym_kernel": ([ctxp, C.POINTER(C.c_double), C.POINTER(C.c_int64)], C.c_int),
        "skb_flow_last_sym_kernel": ([ctxp, C.POINTER(C.c_double), C.POINTER(C.c_int64)], C.c_int),
        "skb_ctx_last_eval_was_symmetric": ([ctxp, C.POINTER(C.c_int)], C.c_int),
        # include/skelly_b200_flow.h
        "skb_flow_create": ([C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_flow_destroy": ([ctxp], C.c_int),
        "skb_flow_set_fibers": ([ctxp, _dp, C.POINTER(C.c_int), _dp, C.c_int], C.c_int),
        "skb_flow_set_periphery": ([ctxp, _dp, _dp, C.c_int64], C.c_int),
        "skb_flow_set_bodies": ([ctxp, _dp, _dp, C.c_int64, _dp, C.c_int], C.c_int),
        "skb_flow_fibers": ([ctxp, _dp, C.c_int64, _dp, C.c_double, C.c_int, _dp], C.c_int),
        "skb_flow_periphery": ([ctxp, _dp, C.c_int64, _dp, C.c_double, _dp], C.c_int),
        "skb_flow_bodies": ([ctxp, _dp, C.c_int64, _dp, _dp, C.c_double, _dp], C.c_int),
        "skb_flow_matvec": ([ctxp, _dp, _dp, _dp, _dp, C.c_double, _dp], C.c_int),
        "skb_flow_last_stats": ([ctxp, C.POINTER(FlowStats)], C.c_int),
        "skb_flow_set_point_sources": ([ctxp, _dp, _dp, _dp, C.c_int], C.c_int),
        "skb_flow_set_background": ([ctxp, C.POINTER(C.c_int), _dp, _dp], C.c_int),
        "skb_flow_velocity_at_targets": ([ctxp, _dp, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, _dp], C.c_int),
        "skb_flow_set_target_window": ([ctxp, C.c_int64, C.c_int64], C.c_int),
        "skb_flow_set_self_exclusion": ([ctxp, C.c_int], C.c_int),
        "skb_flow_set_overlap": ([ctxp, C.c_int], C.c_int),
        "skb_mflow_set_overlap": ([ctxp, C.c_int], C.c_int),
        "skb_dense_apply_background_device": ([ctxp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "skb_flow_set_cross": ([ctxp, C.c_int], C.c_int),
        "skb_mflow_set_cross": ([ctxp, C.c_int], C.c_int),
        "skb_flow_group_init": ([ctxp, C.c_int, C.c_int], C.c_int),
        "skb_flow_group_export": ([ctxp, C.c_void_p], C.c_int),
        "skb_flow_group_import": ([ctxp, C.c_int, C.c_void_p], C.c_int),
        "skb_flow_group_connect": ([ctxp, C.c_int, ctxp], C.c_int),
        "skb_flow_group_error": ([ctxp, C.POINTER(C.c_int)], C.c_int),
        "skb_flow_group_warmup": ([ctxp], C.c_int),
        "skb_flow_group_set_solo": ([ctxp, C.c_int], C.c_int),
        "skb_flow_apply_matvec_device": ([ctxp, ctxp] + [C.c_void_p] * 6 + [C.c_double] + [C.c_void_p] * 4, C.c_int),
        "skb_partition_query": ([C.POINTER(C.c_int), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                 C.POINTER(C.c_int64)], C.c_int),
        "skb_mflow_create": ([C.POINTER(C.c_int), C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_mflow_destroy": ([ctxp], C.c_int),
        "skb_mflow_n_devices": ([ctxp, C.POINTER(C.c_int)], C.c_int),
        "skb_mflow_set_fibers": ([ctxp, _dp, C.POINTER(C.c_int), _dp, C.c_int], C.c_int),
        "skb_mflow_set_periphery": ([ctxp, _dp, _dp, C.c_int64], C.c_int),
        "skb_mflow_set_bodies": ([ctxp, _dp, _dp, C.c_int64, _dp, C.c_int], C.c_int),
        "skb_mflow_set_self_exclusion": ([ctxp, C.c_int], C.c_int),
        "skb_mflow_partition": ([ctxp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.POINTER(C.c_int64)] * 4,
                                C.c_int),
        "skb_mflow_set_fiber_class": ([ctxp, C.c_int, _dp, _dp], C.c_int),
        "skb_mflow_set_fiber_operators": ([ctxp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)], C.c_int),
        "skb_mflow_set_fiber_preconditioner": ([ctxp, _dp], C.c_int),
        "skb_mflow_set_dense": ([ctxp, C.c_int, _dp, C.c_int64, C.c_int64], C.c_int),
        "skb_mflow_matvec": ([ctxp, _dp, _dp, _dp, _dp, C.c_double, _dp], C.c_int),
        "skb_mflow_apply_matvec": ([ctxp, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp], C.c_int),
        "skb_mflow_velocity_at_targets": ([ctxp, _dp, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, _dp], C.c_int),
        "skb_mflow_last_stats": ([ctxp, C.POINTER(FlowStats)], C.c_int),
        "skb_flow_set_target_ranges": ([ctxp, C.c_int, C.c_int] + [C.c_int64] * 4, C.c_int),
        "skb_flow_apply_fiber_force_device": ([ctxp, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "skb_flow_fiber_matvec_device": ([ctxp] + [C.c_void_p] * 5, C.c_int),
        "skb_flow_set_fiber_class": ([ctxp, C.c_int, _dp, _dp], C.c_int),
        "skb_flow_set_fiber_operators": ([ctxp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)], C.c_int),
        "skb_flow_apply_fiber_force": ([ctxp, _dp, _dp], C.c_int),
        "skb_flow_set_fiber_preconditioner": ([ctxp, _dp], C.c_int),
        "skb_flow_apply_fiber_preconditioner": ([ctxp, _dp, _dp], C.c_int),
        "skb_flow_apply_fiber_preconditioner_device": ([ctxp, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "skb_flow_fiber_matvec": ([ctxp, _dp, _dp, _dp, _dp], C.c_int),
        "skb_flow_apply_matvec": ([ctxp, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp], C.c_int),
        "skb_flow_apply_matvec_dense": ([ctxp, ctxp, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp], C.c_int),
        "skb_dense_apply_device": ([ctxp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "skb_dense_shape": ([ctxp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)], C.c_int),
        # include/skelly_b200_dense.h
        "skb_dense_create": ([C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_dense_create_on": ([C.POINTER(C.c_int), C.c_int, C.POINTER(ctxp)], C.c_int),
        "skb_dense_device": ([ctxp, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "skb_dense_destroy": ([ctxp], C.c_int),
        "skb_dense_set_matrix": ([ctxp, C.c_int, _dp, C.c_int64, C.c_int64], C.c_int),
        "skb_dense_apply": ([ctxp, C.c_int, _dp, _dp, _dp], C.c_int),
        "skb_dense_last_stats": ([ctxp, C.POINTER(DenseStats)], C.c_int),
        "skb_flow_matvec_device": ([ctxp] + [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_void_p], C.c_int),
    }
    global BOUND_FUNCTIONS
    BOUND_FUNCTIONS = tuple(sorted(sig))
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise SkbError(f"skelly_b200 error {rc}: {library().skb_last_error_string().decode()}")


def _arr(x, cols):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.size == 0:
        return x.reshape(0, cols)
    if x.ndim != 2 or x.shape[1] != cols:
        raise ValueError(f"expected an (n, {cols}) array, got {x.shape}")
    return x


def _p(x):
    return x.ctypes.data_as(_dp)


def device_count() -> int:
    n = C.c_int(0)
    library().skb_device_count(C.byref(n))
    return n.value


def launch_count() -> int:
    return int(library().skb_launch_count())


def stokeslet_direct(r_src, f_src, r_trg):
    """kernels::stokeslet_direct_gpu_impl semantics (kernels.hpp:17): 1/(8 pi) included, 1/eta not."""
    r_src, f_src, r_trg = _arr(r_src, 3), _arr(f_src, 3), _arr(r_trg, 3)
    if r_src.shape[0] != f_src.shape[0]:
        raise ValueError("r_src / f_src size mismatch")
    u = np.empty((r_trg.shape[0], 3))
    _check(library().skb_stokeslet_direct(_p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0]))
    return u


def stresslet_direct(r_src, f_src, r_trg):
    """kernels::stresslet_direct_gpu_impl semantics (kernels.hpp:19)."""
    r_src, f_src, r_trg = _arr(r_src, 3), _arr(f_src, 9), _arr(r_trg, 3)
    if r_src.shape[0] != f_src.shape[0]:
        raise ValueError("r_src / f_src size mismatch")
    u = np.empty((r_trg.shape[0], 3))
    _check(library().skb_stresslet_direct(_p(r_src), _p(f_src), r_src.shape[0], _p(r_trg), _p(u), r_trg.shape[0]))
    return u


class Context:
    """Evaluator context: positions cached on the device(s), strengths shipped per evaluation."""

    def __init__(self, n_gpus: int = 1, device_ids=None):
        self._h = C.c_void_p()
        L = library()
        if device_ids is None:
            _check(L.skb_ctx_create(int(n_gpus), C.byref(self._h)))
        else:
            ids = (C.c_int * len(device_ids))(*device_ids)
            _check(L.skb_ctx_create_on(ids, len(device_ids), C.byref(self._h)))
        self.n_trg = 0
        self.n_src = {KERNEL_STOKESLET: 0, KERNEL_STRESSLET: 0}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            library().skb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_gpus(self) -> int:
        n = C.c_int(0)
        _check(library().skb_ctx_n_gpus(self._h, C.byref(n)))
        return n.value

    # ---- host-pointer API ----
    def set_targets(self, r_trg):
        r_trg = _arr(r_trg, 3)
        _check(library().skb_set_targets(self._h, _p(r_trg), r_trg.shape[0]))
        self.n_trg = r_trg.shape[0]

    def set_sources(self, kind, r_src):
        r_src = _arr(r_src, 3)
        _check(library().skb_set_sources(self._h, int(kind), _p(r_src), r_src.shape[0]))
        self.n_src[int(kind)] = r_src.shape[0]

    def set_source_exclusion_ids(self, ids):
        """Opt-in fused self-exclusion: Stokeslet sources that are the leading targets; pairs with equal ids (fiber
        indices) contribute exactly 0.  None switches it off."""
        if ids is None:
            _check(library().skb_set_source_exclusion_ids(self._h, None, 0))
            return
        ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1)
        _check(library().skb_set_source_exclusion_ids(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), ids.shape[0]))

    def set_source_normals(self, normals):
        normals = _arr(normals, 3)
        _check(library().skb_set_source_normals(self._h, _p(normals), normals.shape[0]))

    def eval(self, kind, f_src, out=None, accumulate=False):
        f_src = _arr(f_src, 3 if kind == KERNEL_STOKESLET else 9)
        if f_src.shape[0] != self.n_src[int(kind)]:
            raise ValueError("strength count does not match the sources set")
        if out is None:
            if accumulate:
                raise ValueError("accumulate needs an output array")
            out = np.empty((self.n_trg, 3))
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape == (self.n_trg, 3)
        _check(library().skb_eval(self._h, int(kind), _p(f_src), _p(out), int(bool(accumulate))))
        return out

    def eval_fused(self, f_sl=None, f_dl=None):
        out = np.empty((self.n_trg, 3))
        a = _arr(f_sl, 3) if f_sl is not None else None
        b = _arr(f_dl, 9) if f_dl is not None else None
        _check(library().skb_eval_fused(self._h, _p(a) if a is not None else None,
                                        _p(b) if b is not None else None, _p(out)))
        return out

    def eval_double_layer(self, density, eta, out=None, accumulate=False):
        density = _arr(density, 3)
        if out is None:
            out = np.empty((self.n_trg, 3))
        _check(library().skb_eval_double_layer(self._h, _p(density), float(eta), _p(out), int(bool(accumulate))))
        return out

    # ---- device-pointer API (addresses as ints, e.g. torch.Tensor.data_ptr()) ----
    def set_targets_device(self, ptr: int, n_trg: int, stream: int = 0):
        _check(library().skb_set_targets_device(self._h, C.c_void_p(ptr), int(n_trg), C.c_void_p(stream)))
        self.n_trg = int(n_trg)

    def set_sources_device(self, kind, ptr: int, n_src: int, stream: int = 0):
        _check(library().skb_set_sources_device(self._h, int(kind), C.c_void_p(ptr), int(n_src), C.c_void_p(stream)))
        self.n_src[int(kind)] = int(n_src)

    def eval_device(self, kind, f_ptr: int, u_ptr: int, accumulate=False, stream: int = 0):
        _check(library().skb_eval_device(self._h, int(kind), C.c_void_p(f_ptr), C.c_void_p(u_ptr),
                                         int(bool(accumulate)), C.c_void_p(stream)))

    def sync(self):
        _check(library().skb_sync(self._h))

    # ---- instrumentation ----
    def stats(self) -> dict:
        s = EvalStats()
        _check(library().skb_last_eval_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in EvalStats._fields_}

    def set_tuning(self, targets_per_thread=0, source_splits=0):
        _check(library().skb_ctx_set_tuning(self._h, int(targets_per_thread), int(source_splits)))

    def set_symmetric(self, mode: int):
        """-1 auto, 0 off, 1 on: Newton's-third-law kernel for sources that are the leading targets."""
        _check(library().skb_ctx_set_symmetric(self._h, int(mode)))

    def set_sym_partition(self, part: int, n_parts: int):
        """One rank per GPU: evaluate only the block rows of the self-interaction owned by `part`; the leading
        n_src rows of the result are partial sums to be all-reduced over the parts."""
        _check(library().skb_ctx_set_sym_partition(self._h, int(part), int(n_parts)))

    def last_sym_kernel(self):
        ms, pairs = C.c_double(0), C.c_int64(0)
        _check(library().skb_ctx_last_sym_kernel(self._h, C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    def last_eval_was_symmetric(self) -> bool:
        y = C.c_int(0)
        _check(library().skb_ctx_last_eval_was_symmetric(self._h, C.byref(y)))
        return bool(y.value)

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0)
        _check(library().skb_measure_fp64_peak(self._h, C.byref(v)))
        return v.value


class Flow:
    """Device-resident flow() layer (include/skelly_b200_flow.h): FiberContainer / Periphery / BodyContainer flows
    and the fused hydrodynamic part of System::apply_matvec."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(library().skb_flow_create(int(device), C.byref(self._h)))
        self.n_fib = self.n_shell = self.n_body = self.n_bodies = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            library().skb_flow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_fibers(self, r_fib, n_nodes, lengths):
        r_fib = _arr(r_fib, 3)
        n_nodes = np.ascontiguousarray(n_nodes, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.float64)
        assert n_nodes.shape == lengths.shape and int(n_nodes.sum()) == r_fib.shape[0]
        _check(library().skb_flow_set_fibers(self._h, _p(r_fib), n_nodes.ctypes.data_as(C.POINTER(C.c_int)),
                                             _p(lengths), int(n_nodes.shape[0])))
        self.n_fib = r_fib.shape[0]
        self._fiber_off = np.concatenate([[0], np.cumsum(n_nodes, dtype=np.int64)])

    def set_periphery(self, node_pos, node_normal):
        node_pos, node_normal = _arr(node_pos, 3), _arr(node_normal, 3)
        _check(library().skb_flow_set_periphery(self._h, _p(node_pos), _p(node_normal), node_pos.shape[0]))
        self.n_shell = node_pos.shape[0]

    def set_bodies(self, node_pos, node_normal, centers):
        node_pos, node_normal, centers = _arr(node_pos, 3), _arr(node_normal, 3), _arr(centers, 3)
        _check(library().skb_flow_set_bodies(self._h, _p(node_pos), _p(node_normal), node_pos.shape[0], _p(centers),
                                             centers.shape[0]))
        self.n_body, self.n_bodies = node_pos.shape[0], centers.shape[0]

    def fiber_flow(self, r_trg, fib_forces, eta, subtract_self=True):
        r_trg, fib_forces = _arr(r_trg, 3), _arr(fib_forces, 3)
        vel = np.empty((r_trg.shape[0], 3))
        _check(library().skb_flow_fibers(self._h, _p(r_trg), r_trg.shape[0], _p(fib_forces), float(eta),
                                         int(bool(subtract_self)), _p(vel)))
        return vel

    def periphery_flow(self, r_trg, density, eta):
        r_trg, density = _arr(r_trg, 3), _arr(density, 3)
        vel = np.empty((r_trg.shape[0], 3))
        _check(library().skb_flow_periphery(self._h, _p(r_trg), r_trg.shape[0], _p(density), float(eta), _p(vel)))
        return vel

    def body_flow(self, r_trg, densities, forces_torques, eta):
        r_trg, densities = _arr(r_trg, 3), _arr(densities, 3)
        ft = _arr(forces_torques, 6)
        vel = np.empty((r_trg.shape[0], 3))
        _check(library().skb_flow_bodies(self._h, _p(r_trg), r_trg.shape[0], _p(densities), _p(ft), float(eta),
                                         _p(vel)))
        return vel

    def set_point_sources(self, positions, forces, torques):
        positions, forces, torques = _arr(positions, 3), _arr(forces, 3), _arr(torques, 3)
        _check(library().skb_flow_set_point_sources(self._h, _p(positions), _p(forces), _p(torques),
                                                    positions.shape[0]))

    def set_background(self, components, scale_factor, uniform):
        comp = (C.c_int * 3)(*[int(c) for c in components])
        sc = np.ascontiguousarray(scale_factor, dtype=np.float64)
        un = np.ascontiguousarray(uniform, dtype=np.float64)
        _check(library().skb_flow_set_background(self._h, comp, _p(sc), _p(un)))

    def velocity_at_targets(self, r_trg, fib_forces, shell_density, body_densities, body_forces_torques, eta):
        """System::velocity_at_targets (system.cpp:355-359): fiber (no self subtraction) + body + periphery flows."""
        r_trg = _arr(r_trg, 3)
        a, b, c = _arr(fib_forces, 3), _arr(shell_density, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        vel = np.empty((r_trg.shape[0], 3))
        _check(library().skb_flow_velocity_at_targets(self._h, _p(r_trg), r_trg.shape[0], _p(a), _p(b), _p(c), _p(ft),
                                                      float(eta), _p(vel)))
        return vel

    def set_cross(self, mode: int):
        """Fiber <-> periphery pairs of the matvec in one geometry pass: -1 auto (default), 0 never, 1 whenever applicable."""
        _check(library().skb_flow_set_cross(self._h, int(mode)))

    def set_self_exclusion(self, fused: bool):
        """Matvec self term: False = the reference's compute-then-subtract (default); True = the pair kernels skip
        intra-fiber pairs."""
        _check(library().skb_flow_set_self_exclusion(self._h, int(bool(fused))))

    def set_overlap(self, on: bool):
        """Periphery dense operator beside the pair kernels on a side stream (default) or as one kernel at the end."""
        _check(library().skb_flow_set_overlap(self._h, int(bool(on))))

    def set_target_window(self, begin: int, end: int = -1):
        """Evaluate only rows [begin, end) of [fibers | periphery | bodies] in matvec() (one rank's block)."""
        _check(library().skb_flow_set_target_window(self._h, int(begin), int(end)))
        self._window, self._ranges = (int(begin), int(end)), None

    def set_target_ranges(self, fiber_begin, fiber_end, shell_begin, shell_end, body_begin, body_end):
        """The reference's MPI decomposition: whole fibers [fiber_begin, fiber_end), periphery rows, body-node rows;
        matvec() then returns [own fiber nodes | own shell rows | own body rows]."""
        r = tuple(int(v) for v in (fiber_begin, fiber_end, shell_begin, shell_end, body_begin, body_end))
        _check(library().skb_flow_set_target_ranges(self._h, *r))
        self._ranges, self._window = r, None

    def _pieces(self):
        """(own fiber nodes, own shell rows, own body rows) of the current window / ranges."""
        nf, ns, nb = self.n_fib, self.n_shell, self.n_body
        cl = lambda x, lo, hi: min(max(x, lo), hi)
        if getattr(self, "_ranges", None):
            f0, f1, s0, s1, b0, b1 = self._ranges
            off = getattr(self, "_fiber_off", np.zeros(1, dtype=np.int64))
            nfib = len(off) - 1
            f0 = cl(f0, 0, nfib)
            f1 = cl(f1, f0, nfib)
            s0 = cl(s0, 0, ns)
            b0 = cl(b0, 0, nb)
            return int(off[f1] - off[f0]), cl(s1, s0, ns) - s0, cl(b1, b0, nb) - b0
        w0, w1 = getattr(self, "_window", None) or (0, -1)
        n_all = nf + ns + nb
        w0 = cl(w0, 0, n_all)
        w1 = n_all if w1 < 0 else cl(w1, w0, n_all)
        return (cl(w1, 0, nf) - cl(w0, 0, nf), cl(w1 - nf, 0, ns) - cl(w0 - nf, 0, ns),
                cl(w1 - nf - ns, 0, nb) - cl(w0 - nf - ns, 0, nb))

    def matvec_device(self, d_fib_forces: int, d_shell_density: int, d_body_densities: int, d_body_forces: int,
                      d_body_torques: int, eta: float, d_v_window: int, stream: int = 0):
        """Device-pointer matvec (addresses as ints); asynchronous on `stream`."""
        _check(library().skb_flow_matvec_device(self._h, C.c_void_p(d_fib_forces), C.c_void_p(d_shell_density),
                                                C.c_void_p(d_body_densities), C.c_void_p(d_body_forces),
                                                C.c_void_p(d_body_torques), float(eta), C.c_void_p(d_v_window),
                                                C.c_void_p(stream)))

    def matvec(self, fib_forces, shell_density, body_densities, body_forces_torques, eta):
        a, b, c = _arr(fib_forces, 3), _arr(shell_density, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        v = np.empty((sum(self._pieces()), 3))
        _check(library().skb_flow_matvec(self._h, _p(a), _p(b), _p(c), _p(ft), float(eta), _p(v)))
        return v

    # ---- per-fiber dense operators; matrices go in column-major like Eigen's .data() ----
    def set_fiber_class(self, n_nodes: int, D_1_0, P_downsample_bc):
        """FiberFiniteDifference::matrices_.at(n): D_1_0 (n,n), P_downsample_bc (4n-14,4n)."""
        n = int(n_nodes)
        D = np.asfortranarray(D_1_0, dtype=np.float64)
        P = np.asfortranarray(P_downsample_bc, dtype=np.float64)
        if D.shape != (n, n) or P.shape != (4 * n - 14, 4 * n):
            raise ValueError(f"class matrices for n={n}: got {D.shape}, {P.shape}")
        _check(library().skb_flow_set_fiber_class(self._h, n, _p(D), _p(P)))

    @staticmethod
    def _colmajor_concat(mats):
        """Per-fiber matrices -> one buffer, every matrix column-major (Eigen's .data()).  A 3-D C-ordered stack of
        equal shapes is transposed in one pass; a list may be ragged."""
        if isinstance(mats, np.ndarray) and mats.ndim == 3:
            return np.ascontiguousarray(np.asarray(mats, dtype=np.float64).transpose(0, 2, 1)).reshape(-1)
        if len(mats) == 0:
            return np.zeros(0)
        return np.concatenate([np.asarray(m, dtype=np.float64).ravel(order="F") for m in mats])

    def set_fiber_operators(self, A_list, force_list, xs, length_prev, plus_bc_velocity, colmajor=False):
        """A_ (4n,4n) and force_operator_ (3n,4n) per fiber (lists, or (n_fibers, rows, cols) stacks), tangents xs
        (N_f,3), length_prev_, plus-end velocity BC -- of the OWN fibers when target ranges are set.  colmajor=True:
        A_list / force_list are already the concatenated column-major buffers (Eigen's .data()), passed through."""
        if colmajor:
            A = np.ascontiguousarray(A_list, dtype=np.float64).reshape(-1)
            F = np.ascontiguousarray(force_list, dtype=np.float64).reshape(-1)
        else:
            A, F = self._colmajor_concat(A_list), self._colmajor_concat(force_list)
        xs = _arr(xs, 3)
        lp = np.ascontiguousarray(length_prev, dtype=np.float64)
        pl = np.ascontiguousarray(plus_bc_velocity, dtype=np.int32)
        assert xs.shape[0] == self._pieces()[0] and lp.shape == pl.shape and (colmajor or lp.shape == (len(A_list),))
        _check(library().skb_flow_set_fiber_operators(self._h, _p(A), _p(F), _p(xs), _p(lp),
                                                      pl.ctypes.data_as(C.POINTER(C.c_int))))

    def set_fiber_preconditioner(self, A_inv_list):
        """Explicit inverses of A_ (fib.A_LU_.inverse()), same fibers and order as set_fiber_operators."""
        Ai = self._colmajor_concat(A_inv_list)
        _check(library().skb_flow_set_fiber_preconditioner(self._h, _p(Ai)))

    def apply_fiber_preconditioner(self, x_fibers):
        """fc.apply_preconditioner (fcfd.cpp:331-339): y = A_^-1 x per (own) fiber."""
        x = np.ascontiguousarray(x_fibers, dtype=np.float64).reshape(-1)
        assert x.shape[0] == 4 * self._pieces()[0]
        y = np.empty_like(x)
        _check(library().skb_flow_apply_fiber_preconditioner(self._h, _p(x), _p(y)))
        return y

    def apply_fiber_preconditioner_device(self, d_x_fibers: int, d_y: int, stream: int = 0):
        _check(library().skb_flow_apply_fiber_preconditioner_device(self._h, C.c_void_p(d_x_fibers), C.c_void_p(d_y),
                                                                    C.c_void_p(stream)))

    def apply_fiber_force(self, x_fibers):
        """fc.apply_fiber_force (fcfd.cpp:272-287): (4 N_f,) -> fw (N_f, 3)."""
        x = np.ascontiguousarray(x_fibers, dtype=np.float64).reshape(-1)
        n_own = self._pieces()[0]
        assert x.shape[0] == 4 * n_own
        fw = np.empty((n_own, 3))
        _check(library().skb_flow_apply_fiber_force(self._h, _p(x), _p(fw)))
        return fw

    def fiber_matvec(self, x_fibers, v_fibers, v_fib_boundary=None):
        """fc.matvec (fcfd.cpp:216-232): x (4 N_f,), v (N_f,3), v_fib_boundary (n_fibers,7) or None -> res (4 N_f,)."""
        x = np.ascontiguousarray(x_fibers, dtype=np.float64).reshape(-1)
        v = _arr(v_fibers, 3)
        n_own = self._pieces()[0]
        assert x.shape[0] == 4 * n_own and v.shape[0] == n_own
        vb = None if v_fib_boundary is None else _arr(v_fib_boundary, 7)
        res = np.empty(4 * n_own)
        _check(library().skb_flow_fiber_matvec(self._h, _p(x), _p(v), None if vb is None else _p(vb), _p(res)))
        return res

    def apply_fiber_force_device(self, d_x_fibers: int, d_fw: int, stream: int = 0):
        _check(library().skb_flow_apply_fiber_force_device(self._h, C.c_void_p(d_x_fibers), C.c_void_p(d_fw),
                                                           C.c_void_p(stream)))

    def fiber_matvec_device(self, d_x_fibers: int, d_v_fibers: int, d_v_fib_boundary: int, d_res: int,
                            stream: int = 0):
        _check(library().skb_flow_fiber_matvec_device(self._h, C.c_void_p(d_x_fibers), C.c_void_p(d_v_fibers),
                                                      C.c_void_p(d_v_fib_boundary) if d_v_fib_boundary else None,
                                                      C.c_void_p(d_res), C.c_void_p(stream)))

    def apply_matvec(self, x_fibers, shell_density, body_densities, body_forces_torques, eta,
                     fiber_link_conditions=None, dense=None, out=None):
        """System::apply_matvec (system.cpp:298-318) with the fiber operators on the device.
        Returns (res_fibers (4 N_f,), v_shell (N_s,3), v_bodies (N_b,3)); with dense= a single-device Dense holding
        stresslet_plus_complementary the second item is res_shell = shell.matvec(x_shell, v_shell) instead."""
        x = np.ascontiguousarray(x_fibers, dtype=np.float64).reshape(-1)
        assert x.shape[0] == 4 * self.n_fib
        b, c = _arr(shell_density, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        vb = None if fiber_link_conditions is None else _arr(fiber_link_conditions, 7)
        if out is not None:  # caller's (e.g. pinned) buffers: (res (4 N_f,), v_shell (N_s,3), v_bodies (N_b,3))
            res, v_s, v_b = out
            assert res.shape == (4 * self.n_fib,) and v_s.shape == (self.n_shell, 3) and v_b.shape == (self.n_body, 3)
            assert all(a.flags.c_contiguous and a.dtype == np.float64 for a in out)
        else:
            res = np.empty(4 * self.n_fib)
            v_s, v_b = np.empty((self.n_shell, 3)), np.empty((self.n_body, 3))
        if dense is not None:
            _check(library().skb_flow_apply_matvec_dense(self._h, dense._h, _p(x), _p(b), _p(c), _p(ft),
                                                         None if vb is None else _p(vb), float(eta), _p(res), _p(v_s),
                                                         _p(v_b)))
        else:
            _check(library().skb_flow_apply_matvec(self._h, _p(x), _p(b), _p(c), _p(ft),
                                                   None if vb is None else _p(vb), float(eta), _p(res), _p(v_s),
                                                   _p(v_b)))
        return res, v_s, v_b

    def apply_matvec_device(self, dense, d_x_fibers: int, d_x_shell: int, d_body_densities: int, d_body_forces: int,
                            d_body_torques: int, d_fiber_link_conditions: int, eta: float, d_res_fibers: int,
                            d_out_shell: int, d_v_bodies: int, stream: int = 0):
        """System::apply_matvec with every operand already on the device (addresses as ints, 0 = NULL); OWN slices
        for a group member.  Asynchronous on `stream`."""
        vp = lambda a: C.c_void_p(a) if a else None
        _check(library().skb_flow_apply_matvec_device(self._h, dense._h if dense is not None else None, vp(d_x_fibers),
                                                      vp(d_x_shell), vp(d_body_densities), vp(d_body_forces),
                                                      vp(d_body_torques), vp(d_fiber_link_conditions), float(eta),
                                                      vp(d_res_fibers), vp(d_out_shell), vp(d_v_bodies), vp(stream)))

    # ---- multi-GPU groups (peer memory) ----
    def last_sym_kernel(self):
        """(ms, ordered pairs) of the last launch of the symmetric fiber-fiber kernel inside matvec()."""
        ms, pairs = C.c_double(0), C.c_int64(0)
        _check(library().skb_flow_last_sym_kernel(self._h, C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    def group_init(self, rank: int, size: int):
        _check(library().skb_flow_group_init(self._h, int(rank), int(size)))

    def group_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(library().skb_flow_group_export(self._h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def group_import(self, peer_rank: int, handle: bytes):
        assert len(handle) == 64
        buf = C.create_string_buffer(handle, 64)
        _check(library().skb_flow_group_import(self._h, int(peer_rank), C.cast(buf, C.c_void_p)))

    def group_connect(self, peer_rank: int, peer: "Flow"):
        _check(library().skb_flow_group_connect(self._h, int(peer_rank), peer._h))

    def group_warmup(self):
        _check(library().skb_flow_group_warmup(self._h))

    def group_set_solo(self, solo: bool):
        _check(library().skb_flow_group_set_solo(self._h, int(bool(solo))))

    def group_error(self) -> int:
        m = C.c_int(-1)
        _check(library().skb_flow_group_error(self._h, C.byref(m)))
        return m.value

    def stats(self) -> dict:
        s = FlowStats()
        _check(library().skb_flow_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in FlowStats._fields_}


class MultiFlow:
    """skb_mflow: ONE process driving n GPUs (include/skelly_b200_flow.h) -- same conventions as Flow, complete host
    arrays in and out; fibers / periphery rows / body rows are partitioned over the devices, which exchange strengths
    and partial velocities through peer memory."""

    def __init__(self, devices):
        devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = (C.c_int * len(devs))(*devs)
        self._h = C.c_void_p()
        _check(library().skb_mflow_create(arr, len(devs), C.byref(self._h)))
        self.devices = devs
        self.n_fib = self.n_shell = self.n_body = self.n_bodies = self.n_fibers = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            library().skb_mflow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_fibers(self, r_fib, n_nodes, lengths):
        r_fib = _arr(r_fib, 3)
        n_nodes = np.ascontiguousarray(n_nodes, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.float64)
        assert n_nodes.shape == lengths.shape and int(n_nodes.sum()) == r_fib.shape[0]
        _check(library().skb_mflow_set_fibers(self._h, _p(r_fib), n_nodes.ctypes.data_as(C.POINTER(C.c_int)),
                                              _p(lengths), int(n_nodes.shape[0])))
        self.n_fib, self.n_fibers = r_fib.shape[0], int(n_nodes.shape[0])

    def set_periphery(self, node_pos, node_normal):
        node_pos, node_normal = _arr(node_pos, 3), _arr(node_normal, 3)
        _check(library().skb_mflow_set_periphery(self._h, _p(node_pos), _p(node_normal), node_pos.shape[0]))
        self.n_shell = node_pos.shape[0]

    def set_bodies(self, node_pos, node_normal, centers):
        node_pos, node_normal, centers = _arr(node_pos, 3), _arr(node_normal, 3), _arr(centers, 3)
        _check(library().skb_mflow_set_bodies(self._h, _p(node_pos), _p(node_normal), node_pos.shape[0], _p(centers),
                                              centers.shape[0]))
        self.n_body, self.n_bodies = node_pos.shape[0], centers.shape[0]

    def set_self_exclusion(self, fused: bool):
        _check(library().skb_mflow_set_self_exclusion(self._h, int(bool(fused))))

    def set_cross(self, mode: int):
        _check(library().skb_mflow_set_cross(self._h, int(mode)))

    def set_overlap(self, on: bool):
        _check(library().skb_mflow_set_overlap(self._h, int(bool(on))))

    def partition(self, member: int):
        f0, f1 = C.c_int(), C.c_int()
        r = [C.c_int64() for _ in range(4)]
        _check(library().skb_mflow_partition(self._h, int(member), C.byref(f0), C.byref(f1), *[C.byref(x) for x in r]))
        return (f0.value, f1.value) + tuple(x.value for x in r)

    def set_fiber_class(self, n_nodes: int, D_1_0, P_downsample_bc):
        n = int(n_nodes)
        D = np.asfortranarray(D_1_0, dtype=np.float64)
        P = np.asfortranarray(P_downsample_bc, dtype=np.float64)
        if D.shape != (n, n) or P.shape != (4 * n - 14, 4 * n):
            raise ValueError(f"class matrices for n={n}: got {D.shape}, {P.shape}")
        _check(library().skb_mflow_set_fiber_class(self._h, n, _p(D), _p(P)))

    def set_fiber_operators(self, A_list, force_list, xs, length_prev, plus_bc_velocity, colmajor=False):
        if colmajor:
            A = np.ascontiguousarray(A_list, dtype=np.float64).reshape(-1)
            F = np.ascontiguousarray(force_list, dtype=np.float64).reshape(-1)
        else:
            A, F = Flow._colmajor_concat(A_list), Flow._colmajor_concat(force_list)
        xs = _arr(xs, 3)
        lp = np.ascontiguousarray(length_prev, dtype=np.float64)
        pl = np.ascontiguousarray(plus_bc_velocity, dtype=np.int32)
        assert xs.shape[0] == self.n_fib and lp.shape == pl.shape == (self.n_fibers,)
        _check(library().skb_mflow_set_fiber_operators(self._h, _p(A), _p(F), _p(xs), _p(lp),
                                                       pl.ctypes.data_as(C.POINTER(C.c_int))))

    def set_fiber_preconditioner(self, A_inv_list):
        Ai = Flow._colmajor_concat(A_inv_list)
        _check(library().skb_mflow_set_fiber_preconditioner(self._h, _p(Ai)))

    def set_dense(self, op: int, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        _check(library().skb_mflow_set_dense(self._h, int(op), _p(A), A.shape[0], A.shape[1]))
        self._has_dense = True

    def matvec(self, fib_forces, shell_density, body_densities, body_forces_torques, eta):
        a, b, c = _arr(fib_forces, 3), _arr(shell_density, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        v = np.empty((self.n_fib + self.n_shell + self.n_body, 3))
        _check(library().skb_mflow_matvec(self._h, _p(a), _p(b), _p(c), _p(ft), float(eta), _p(v)))
        return v

    def apply_matvec(self, x_fibers, x_shell, body_densities, body_forces_torques, eta, fiber_link_conditions=None):
        """Returns (res_fibers (4 N_f,), out_shell (N_s,3), v_bodies (N_b,3)); out_shell is res_shell once
        set_dense(DENSE_STRESSLET_PLUS_COMPLEMENTARY, ...) was called, else v_shell."""
        x = np.ascontiguousarray(x_fibers, dtype=np.float64).reshape(-1)
        assert x.shape[0] == 4 * self.n_fib
        b, c = _arr(x_shell, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        vb = None if fiber_link_conditions is None else _arr(fiber_link_conditions, 7)
        res = np.empty(4 * self.n_fib)
        o_s, v_b = np.empty((self.n_shell, 3)), np.empty((self.n_body, 3))
        _check(library().skb_mflow_apply_matvec(self._h, _p(x), _p(b), _p(c), _p(ft), None if vb is None else _p(vb),
                                                float(eta), _p(res), _p(o_s), _p(v_b)))
        return res, o_s, v_b

    def velocity_at_targets(self, r_trg, fib_forces, shell_density, body_densities, body_forces_torques, eta):
        r_trg = _arr(r_trg, 3)
        a, b, c = _arr(fib_forces, 3), _arr(shell_density, 3), _arr(body_densities, 3)
        ft = _arr(body_forces_torques, 6)
        vel = np.empty((r_trg.shape[0], 3))
        _check(library().skb_mflow_velocity_at_targets(self._h, _p(r_trg), r_trg.shape[0], _p(a), _p(b), _p(c), _p(ft),
                                                       float(eta), _p(vel)))
        return vel

    def stats(self) -> dict:
        s = FlowStats()
        _check(library().skb_mflow_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in FlowStats._fields_}


DENSE_STRESSLET_PLUS_COMPLEMENTARY = 0
DENSE_M_INV = 1


class Dense:
    """The periphery's dense operators on the GPU (include/skelly_b200_dense.h): Periphery::matvec and
    Periphery::apply_preconditioner (periphery.cpp:21-47) as row-partitioned GEMVs."""

    def __init__(self, n_gpus: int = 1, device_ids=None):
        self._h = C.c_void_p()
        if device_ids is not None:
            arr = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
            _check(library().skb_dense_create_on(arr, len(device_ids), C.byref(self._h)))
        else:
            _check(library().skb_dense_create(int(n_gpus), C.byref(self._h)))
        self.shape = {}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            library().skb_dense_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_matrix(self, op: int, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        assert A.ndim == 2
        _check(library().skb_dense_set_matrix(self._h, int(op), _p(A), A.shape[0], A.shape[1]))
        self.shape[int(op)] = A.shape

    def apply(self, op: int, x, v_add=None):
        rows, cols = self.shape[int(op)]
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert x.shape[0] == cols
        y = np.empty(rows)
        v = None
        if v_add is not None:
            v = np.ascontiguousarray(v_add, dtype=np.float64).reshape(-1)
            assert v.shape[0] == rows
        _check(library().skb_dense_apply(self._h, int(op), _p(x), _p(v) if v is not None else None, _p(y)))
        return y

    def apply_device(self, op: int, d_x: int, d_v_add: int, d_y: int, stream: int = 0):
        """Single-device handle, operands already on its device (addresses as ints, 0 = NULL v_add); asynchronous."""
        _check(library().skb_dense_apply_device(self._h, int(op), C.c_void_p(d_x), C.c_void_p(d_v_add) if d_v_add else None,
                                                C.c_void_p(d_y), C.c_void_p(stream) if stream else None))

    def apply_background_device(self, op: int, d_x: int, d_y: int, stream: int = 0):
        """y = A x with the background row streamer (one small CTA per SM, TMA-fed): the form that can share the SMs
        with the pair kernels.  Single-device handle, device addresses as ints; asynchronous."""
        _check(library().skb_dense_apply_background_device(self._h, int(op), C.c_void_p(d_x), C.c_void_p(d_y),
                                                           C.c_void_p(stream) if stream else None))

    def stats(self) -> dict:
        s = DenseStats()
        _check(library().skb_dense_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in DenseStats._fields_}


def plan_query(kind, n_trg, n_src, num_sms=148, occupancy=(8, 6, 4, 2), force_T=0, force_S=0) -> dict:
    """Host-side launch planner of the plain pair kernel (no GPU needed)."""
    occ = (C.c_int * 4)(*occupancy)
    T, S, per, gx = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(library().skb_plan_query(int(kind), int(n_trg), int(n_src), int(num_sms), occ, int(force_T), int(force_S),
                                    C.byref(T), C.byref(S), C.byref(per), C.byref(gx)))
    return {"T": T.value, "n_splits": S.value, "tiles_per_split": per.value, "grid_x": gx.value}


def partition_query(fiber_n_nodes, n_shell, n_body, n_members, member):
    """Rows of [fibers | periphery | bodies] member `member` of `n_members` owns (the partition of skb_mflow; no GPU
    needed): (fiber_begin, fiber_end, shell_begin, shell_end, body_begin, body_end)."""
    nn = np.ascontiguousarray(fiber_n_nodes, dtype=np.int32)
    out = (C.c_int64 * 6)()
    _check(library().skb_partition_query(nn.ctypes.data_as(C.POINTER(C.c_int)), int(nn.shape[0]), int(n_shell),
                                         int(n_body), int(n_members), int(member), out))
    return tuple(int(v) for v in out)


def sym_groups_per_block() -> int:
    return int(library().skb_sym_groups_per_block())


def sym_plan_query(n_blocks, part=0, n_parts=1, num_sms=148):
    """Host-side work list of the symmetric kernel (no GPU needed): list of (I, g0, g1, slot), row_begin; block I
    meets the 32-node groups [g0, g1) (sym_groups_per_block() groups per block)."""
    n = C.c_int()
    _check(library().skb_sym_plan_query(int(n_blocks), int(part), int(n_parts), int(num_sms), 0, None, C.byref(n),
                                        None))
    items = (C.c_int * (4 * max(n.value, 1)))()
    rb = (C.c_int * (n_blocks + 1))()
    _check(library().skb_sym_plan_query(int(n_blocks), int(part), int(n_parts), int(num_sms), n.value, items,
                                        C.byref(n), rb))
    return [tuple(items[4 * i:4 * i + 4]) for i in range(n.value)], list(rb)
