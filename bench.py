#!/usr/bin/env python
"""bench.py -- headline benchmark of the SkellySim pair-kernel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): Stokeslet pair-interactions/s.  One "step" = one Stokeslet evaluator call of a GMRES
matvec (FiberContainer::flow's all-pairs call, fiber nodes -> fiber+shell nodes) on the BASELINE `configs[1]`
geometry (ellipsoidal periphery + 1000 fibers x 32 nodes), positions resident on the device(s) exactly as they are
between the matvecs of one timestep, strengths changing every step.  Every ordered (source, target) pair counts once;
the fiber-fiber block is evaluated with the Newton's-third-law kernel (both directions from one geometry pass).

  value : whole-job pairs/s with the step's strengths already in HBM (device-pointer C-ABI entry points).
  e2e   : N = 1: the call a SkellySim evaluator makes -- `skb_eval` with HOST buffers (pinned), H2D of the strengths
          and D2H of the velocities inside the call, wall clock.  N > 1: pinned H2D + collectives + eval + D2H.
  N > 1 : one rank per GPU, weak scaling (the suspension grows so that pairs per GPU stay fixed: nodes ~ sqrt(N)).
          Every rank owns a serpentine set of block rows of the fiber-fiber interaction and a block of the remaining
          targets; per step ONE NCCL all-gather of the source strengths and ONE reduce-scatter of the fiber
          velocities (overlapped with the remainder targets).  `--no-symmetric`: plain kernel, targets block-partitioned,
          all-gather only.
  extras: `matvec` = hydrodynamic part of System::apply_matvec at 102 400 nodes (BASELINE C3), strong-scaled over the
          ranks by target windows; `periphery_dense` = the periphery's dense operator (HBM-bound GEMV), N = 1 only;
          `cpu_baseline` = CPU port of kernels::stokeslet_direct_cpu on all host cores (bounded sample), N = 1 only.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port of the reference's OpenMP direct path
(oracle/, all host threads) on a bounded sample of the same workload; ranks > 0 exit at once.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SL_FLOP_PER_PAIR = 28  # SURVEY.md 8d / BASELINE.md 2.1 (kernels.cu:65-75)
DL_FLOP_PER_PAIR = 40
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 DFMA/clk x 2 flop x max SM clock = 37.2


# ------------------------------------------------------------------------------------------------
# synthetic suspension
# ------------------------------------------------------------------------------------------------
def make_suspension(n_fibers: int, n_shell: int, seed: int = 1, n_nodes: int = 32, length: float = 1.0):
    """Straight fibers x0 + L*linspace(0,1,n)*nhat (skelly_config.py:306-308) with random centres inside the
    ellipsoid a,b,c = 7.8,4.16,4.16 (skelly_config.py:548-550) scaled with the fiber count; shell nodes on the
    1.04x surface (precompute.py:34) with inward normals (precompute.py:78-80)."""
    rng = np.random.default_rng(seed)
    scale = max(1.0, (n_fibers / 1000.0) ** (1.0 / 3.0))
    abc = np.array([7.8, 4.16, 4.16]) * scale
    c = rng.normal(size=(n_fibers, 3))
    c /= np.linalg.norm(c, axis=1)[:, None]
    c *= (rng.uniform(0, 1, n_fibers) ** (1 / 3))[:, None] * 0.85
    centres = c * abc
    nh = rng.normal(size=(n_fibers, 3))
    nh /= np.linalg.norm(nh, axis=1)[:, None]
    s = np.linspace(-0.5 * length, 0.5 * length, n_nodes)
    fib = (centres[:, None, :] + s[None, :, None] * nh[:, None, :]).reshape(-1, 3)
    d = rng.normal(size=(n_shell, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    shell = d * abc * 1.04
    nrm = -(shell / (abc * 1.04) ** 2)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    return np.ascontiguousarray(fib), np.ascontiguousarray(shell), np.ascontiguousarray(nrm)


def workload_sizes(name: str, n_gpus: int):
    g = math.sqrt(n_gpus)
    if name == "c2":   # configs[1]: 1000 fibers x 32 + 8000-node ellipsoid shell
        return int(round(1000 * g)), int(round(8000 * g))
    if name == "c3":   # configs[2]-like: 3000 x 32 + 6000 shell nodes (~1e5 nodes)
        return int(round(3000 * g)), int(round(6400 * g))
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region (pynvml, no subprocess)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self._nv = None
            self.err = str(e)
        self.period = period_s

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of kernels::stokeslet_direct_cpu, all host threads)
# ------------------------------------------------------------------------------------------------
_BEST_THREADS = None


def host_threads() -> int:
    """Thread count of the CPU arms: every logical CPU of the affinity mask, or half of them (one per physical core
    on an SMT-2 host) when that is faster -- decided once by cpu_pick_threads(); never OMP_NUM_THREADS, which torchrun
    sets to 1."""
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_pick_threads(orc, r_src, f_src, r_trg):
    """Give the CPU arm its best shot: time a probe with all logical CPUs and with half of them, keep the faster."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    full = host_threads()
    cands = sorted({full, max(1, full // 2)}, reverse=True)
    n = min(r_trg.shape[0], max(512, 16 * full))
    best, best_t = full, float("inf")
    for th in cands:
        orc.stokeslet_direct_cpu(r_src, f_src, r_trg[:n], 1.0, th)
        t0 = time.perf_counter()
        for _ in range(3):
            orc.stokeslet_direct_cpu(r_src, f_src, r_trg[:n], 1.0, th)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
    _BEST_THREADS = best
    return best


def cpu_sample_plan(orc, r_src, f_src, r_trg, budget_s: float):
    """Pick how many targets a bounded CPU sample evaluates so that one call takes ~budget_s."""
    cpu_pick_threads(orc, r_src, f_src, r_trg)
    n_probe = min(r_trg.shape[0], max(256, 16 * host_threads()))
    t0 = time.perf_counter()
    orc.stokeslet_direct_cpu(r_src, f_src, r_trg[:n_probe], 1.0, host_threads())
    t0 = time.perf_counter()
    orc.stokeslet_direct_cpu(r_src, f_src, r_trg[:n_probe], 1.0, host_threads())
    dt = time.perf_counter() - t0
    rate = r_src.shape[0] * n_probe / max(dt, 1e-9)
    n = int(min(r_trg.shape[0], max(n_probe, rate * budget_s / r_src.shape[0])))
    return n, rate


def cpu_baseline_leg(r_src, f_src, r_trg, budget_s=10.0):
    import oracle as orc
    n, _ = cpu_sample_plan(orc, r_src, f_src, r_trg, budget_s)
    calls, t0 = 0, time.perf_counter()
    while True:  # bounded sample: repeat the call until ~budget_s of CPU work has been timed
        orc.stokeslet_direct_cpu(r_src, f_src, r_trg[:n], 1.0, host_threads())
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or calls >= 10000:
            break
    return {"value": calls * r_src.shape[0] * n / dt, "unit": "pairs/s", "cores": host_threads(), "kind": "port",
            "simd": {0: "scalar", 1: "avx2+fma", 2: "avx512"}[orc.simd_level()],
            "sample": f"all {r_src.shape[0]} sources x first {n} of {r_trg.shape[0]} targets, {calls} calls, "
                      f"{dt:.2f} s; OpenMP static target chunks as kernels.cpp:42-65 (the reference CPU path itself "
                      "needs PVFMM: unbuildable here)"}


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port; the reference's own
    kernels.cpp cannot be compiled without PVFMM/Eigen/MPI)."""
    if rank != 0:
        return
    import oracle as orc
    n_fib, n_shell = workload_sizes(args.workload, args.gpus)
    fib, shell, _ = make_suspension(n_fib, n_shell)
    r_src, r_trg = fib, np.concatenate([fib, shell])
    rng = np.random.default_rng(7)
    f = rng.uniform(-1, 1, r_src.shape)
    per_step = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    n, _ = cpu_sample_plan(orc, r_src, f, r_trg, per_step)
    for _ in range(args.warmup):
        orc.stokeslet_direct_cpu(r_src, f, r_trg[:n], 1.0, host_threads())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.stokeslet_direct_cpu(r_src, f, r_trg[:n], 1.0, host_threads())
    dt = time.perf_counter() - t0
    val = args.steps * r_src.shape[0] * n / dt
    sample = f"each step = all {r_src.shape[0]} sources x first {n} of {r_trg.shape[0]} targets"
    print(json.dumps({
        "impl": "reference", "metric": "stokeslet_pair_interactions_per_s", "value": val, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, n_fib, n_shell), "n_src": int(r_src.shape[0]),
                   "n_trg": int(r_trg.shape[0]), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": host_threads(), "kind": "port",
                         "simd": {0: "scalar", 1: "avx2+fma", 2: "avx512"}[orc.simd_level()], "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


_T0 = time.perf_counter()


def _log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        print(f"[bench +{time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def workload_name(w, n_fib, n_shell):
    return (f"{w}: ellipsoid periphery ({n_shell} nodes) + {n_fib} fibers x 32 nodes; Stokeslet call of "
            f"FiberContainer::flow (fiber nodes -> fiber+shell nodes), FP64 direct kernel")


# ------------------------------------------------------------------------------------------------
# second half of BASELINE's metric: GMRES matvec (hydrodynamic part) at N = 1e5 nodes, strong-scaled over the ranks
# ------------------------------------------------------------------------------------------------
def make_c3_system(seed=2):
    """BASELINE C3 / SURVEY.md 8d S4: 3000 fibers x 32 + 6000 shell nodes + 1 body x 400 nodes = 102 400 nodes."""
    fib, shell, nrm = make_suspension(3000, 6000, seed=seed)
    rng = np.random.default_rng(seed)
    e = rng.normal(size=(400, 3))
    e /= np.linalg.norm(e, axis=1)[:, None]
    centers = np.zeros((1, 3))
    body_pos = centers + 0.5 * e
    return dict(fib=fib, n_nodes=np.full(3000, 32, dtype=np.int32), lengths=np.ones(3000), shell=shell,
                shell_n=nrm, body=np.ascontiguousarray(body_pos), body_n=np.ascontiguousarray(e), centers=centers,
                ff=rng.uniform(-1, 1, fib.shape), sd=rng.uniform(-1, 1, shell.shape),
                bd=rng.uniform(-1, 1, body_pos.shape), force=rng.uniform(-1, 1, (1, 3)),
                torque=rng.uniform(-1, 1, (1, 3)))


def matvec_leg(torch, dist, skb, dev, local_rank, rank, world, steps=20, warmup=3):
    from skellysim_b200.distributed import allgather_strengths, block_range
    g = make_c3_system()
    nf, ns, nb = g["fib"].shape[0], g["shell"].shape[0], g["body"].shape[0]
    n_all = nf + ns + nb
    eta = 1.0
    fl = skb.Flow(local_rank)
    fl.set_fibers(g["fib"], g["n_nodes"], g["lengths"])
    fl.set_periphery(g["shell"], g["shell_n"])
    fl.set_bodies(g["body"], g["body_n"], g["centers"])
    w0, w1 = block_range(n_all, world, rank)
    fl.set_target_window(w0, w1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    chunk = -(-nf // world)
    d_ff = torch.zeros((chunk * world, 3), dtype=torch.float64, device=dev)
    b, e = block_range(nf, world, rank)
    d_mine = d_ff[rank * chunk:(rank + 1) * chunk]
    d_mine[:e - b] = t(g["ff"][b:e])  # each rank owns the forces of its own fibers
    d_sd, d_bd, d_f, d_t = t(g["sd"]), t(g["bd"]), t(g["force"]), t(g["torque"])
    d_v = torch.empty((max(w1 - w0, 1), 3), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        allgather_strengths(d_ff, d_mine)
        fl.matvec_device(d_ff.data_ptr(), d_sd.data_ptr(), d_bd.data_ptr(), d_f.data_ptr(), d_t.data_ptr(), eta,
                         d_v.data_ptr(), stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    for a, bb in evs:
        flush.zero_()
        a.record()
        step()
        bb.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(bb) for a, bb in evs) / steps
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    out = None
    if rank == 0:
        import oracle as orc
        # accuracy gate on shell rows (fiber flow + body flow act there; the shell's own flow does not, system.cpp:301-315)
        st = fl.stats()
        res = {"workload": "c3: 3000 fibers x 32 + 6000 periphery nodes + 1 body x 400 nodes (102400 nodes); "
                           "hydrodynamic part of System::apply_matvec (SL fibers->all, DL shell->fibers+bodies, "
                           "DL body->all, SL+rotlet centres->all, self-term subtraction)",
               "n_nodes": n_all, "ms": ms, "scaling": "strong", "steps": steps,
               "pairs": float(nf) * n_all + float(ns) * (nf + nb) + float(nb + 2) * n_all,
               "launches_per_matvec_rank0": st["launches"]}
        if world == 1:
            v = d_v.cpu().numpy()
            idx = nf + np.random.default_rng(1).choice(ns, 64, replace=False)
            r_all = np.concatenate([g["fib"], g["shell"], g["body"]])
            w = np.tile(orc.trapezoid_weights(32, 1.0), 3000)
            ref = orc.stokeslet_direct_cpu(g["fib"], g["ff"] * w[:, None], r_all[idx], eta)
            ref += orc.body_flow(r_all[idx], g["body"], g["body_n"], g["bd"], g["centers"], g["force"], g["torque"],
                                 eta)
            res["max_rel_err_vs_oracle_shell_rows"] = float(np.abs(v[idx] - ref).max() / np.abs(ref).max())
            # end to end through the host-pointer C ABI (H2D of all strengths + D2H of v_all inside)
            ft = np.concatenate([g["force"], g["torque"]], axis=1)
            fl.matvec(g["ff"], g["sd"], g["bd"], ft, eta)
            t0 = time.perf_counter()
            for _ in range(5):
                fl.matvec(g["ff"], g["sd"], g["bd"], ft, eta)
            res["e2e_ms"] = 1e3 * (time.perf_counter() - t0) / 5
        out = res
    fl.close()
    return out


# ------------------------------------------------------------------------------------------------
# "next" row N1: the periphery's dense operator (HBM-bound GEMV), N_s = 6000 nodes -> 18000 x 18000 FP64
# ------------------------------------------------------------------------------------------------
def dense_leg(skb, hbm_peak_gbs, n_nodes=6000, reps=10):
    n = 3 * n_nodes
    rng = np.random.default_rng(4)
    _log("dense: generating matrix")
    A = rng.random((n, n))  # 2.6 GB
    _log("dense: matrix ready")
    x = rng.random(n)
    v = rng.random(n)
    with skb.Dense(1) as dn:
        dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, A)
        _log("dense: uploaded")
        y = dn.apply(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, x, v)
        ks, ts = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            y = dn.apply(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, x, v)
            ts.append(1e3 * (time.perf_counter() - t0))
            ks.append(dn.stats()["kernel_ms"])
        bytes_ = dn.stats()["bytes"]
    _log("dense: timed")
    rows = np.random.default_rng(0).choice(n, 32, replace=False)
    ref = A[rows] @ x + v[rows]
    err = float(np.max(np.abs(y[rows] - ref) / (np.abs(A[rows]) @ np.abs(x) + np.abs(v[rows]))))
    k = float(np.median(ks))
    return {"workload": f"Periphery::matvec dense operator, {n_nodes} nodes: {n} x {n} FP64 row-major GEMV (+v)",
            "kernel_ms": k, "e2e_ms": float(np.median(ts)), "bytes": int(bytes_),
            "roofline": {"bound": "hbm", "achieved": bytes_ / (k * 1e-3) / 1e9, "peak": hbm_peak_gbs, "unit": "GB/s",
                         "frac": bytes_ / (k * 1e-3) / 1e9 / hbm_peak_gbs},
            "max_backward_err": err}


# ------------------------------------------------------------------------------------------------
# "next" row N2: per-fiber dense operators on the device -- System::apply_matvec end to end for the fiber rows
# (apply_fiber_force -> flows -> fc.matvec), C3 geometry, random dense stand-ins for A_ / force_operator_
# ------------------------------------------------------------------------------------------------
def fiber_ops_leg(skb, hbm_peak_gbs, reps=10):
    import oracle as orc
    g = make_c3_system()
    nf, ns, nb = g["fib"].shape[0], g["shell"].shape[0], g["body"].shape[0]
    n, n_fibers, eta = 32, 3000, 1.0
    rng = np.random.default_rng(6)
    _log("fiber ops: generating operators")
    A = rng.standard_normal((n_fibers, 4 * n, 4 * n)) / np.sqrt(4 * n)       # 393 MB
    F = rng.standard_normal((n_fibers, 3 * n, 4 * n)) / np.sqrt(4 * n)       # 295 MB
    D = rng.standard_normal((n, n))
    P = rng.standard_normal((4 * n - 14, 4 * n)) / np.sqrt(4 * n)
    xs = rng.standard_normal((nf, 3))
    xs /= np.linalg.norm(xs, axis=1)[:, None]
    lprev, plus = np.ones(n_fibers), rng.integers(0, 2, n_fibers).astype(np.int32)
    x = rng.standard_normal(4 * nf)
    link = rng.standard_normal((n_fibers, 7))
    ft = np.concatenate([g["force"], g["torque"]], axis=1)
    with skb.Flow(0) as fl:
        fl.set_fibers(g["fib"], g["n_nodes"], g["lengths"])
        fl.set_periphery(g["shell"], g["shell_n"])
        fl.set_bodies(g["body"], g["body_n"], g["centers"])
        fl.set_fiber_class(n, D, P)
        t0 = time.perf_counter()
        # (the binding lays every matrix out column-major, as Eigen's .data())
        fl.set_fiber_operators(A, F, xs, lprev, plus)
        set_ms = 1e3 * (time.perf_counter() - t0)
        _log("fiber ops: uploaded")
        res, v_s, v_b = fl.apply_matvec(x, g["sd"], g["bd"], ft, eta, link)
        dev, wall = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            res, v_s, v_b = fl.apply_matvec(x, g["sd"], g["bd"], ft, eta, link)
            wall.append(1e3 * (time.perf_counter() - t0))
            dev.append(fl.stats()["device_ms"])
        launches = fl.stats()["launches"]
        k_force, k_mv = [], []
        v_fib = np.ascontiguousarray(rng.standard_normal((nf, 3)))
        for _ in range(reps):
            fw = fl.apply_fiber_force(x)
            k_force.append(fl.stats()["device_ms"])
            fl.fiber_matvec(x, v_fib, link)
            k_mv.append(fl.stats()["device_ms"])
        # fc.apply_preconditioner as a GEMV over explicit inverses (timing only: A itself stands in for A^-1)
        fl.set_fiber_preconditioner(A)
        k_pc = []
        for _ in range(reps):
            fl.apply_fiber_preconditioner(x)
            k_pc.append(fl.stats()["device_ms"])
        # accuracy: sampled fibers against the oracle, the velocities taken from the (separately gated) flow matvec
        v_all = fl.matvec(fw, g["sd"], g["bd"], ft, eta)
    sel = rng.choice(n_fibers, 16, replace=False)
    e_f = e_r = 0.0
    for i in sel:
        s4, s1 = slice(4 * n * i, 4 * n * (i + 1)), slice(n * i, n * (i + 1))
        ref_fw = orc.apply_fiber_force([F[i]], x[s4], [n])
        e_f = max(e_f, float(np.abs(fw[s1] - ref_fw).max() / np.abs(ref_fw).max()))
        ref = orc.fiber_matvec(A[i], D, P, xs[s1], lprev[i], plus[i], x[s4], v_all[s1], link[i])
        e_r = max(e_r, float(np.abs(res[s4] - ref).max() / np.abs(ref).max()))
    # the host work this replaces, batched BLAS on this box's cores (the reference loops over fibers with Eigen GEMVs)
    xb = x.reshape(n_fibers, 4 * n, 1)
    t0 = time.perf_counter()
    for _ in range(3):
        np.matmul(F, xb)
        np.matmul(A, xb)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / 3
    mat_bytes = (A.nbytes + F.nbytes)
    k = float(np.median(k_force) + np.median(k_mv))
    gbs = mat_bytes / (k * 1e-3) / 1e9
    return {"workload": "c3 (3000 fibers x 32 nodes): System::apply_matvec with apply_fiber_force (3n x 4n) and fc.matvec "
                        "(4n x 4n A_, P_downsample_bc, D_1) as batched device GEMVs; fw and v_fibers never leave the GPU",
            "apply_matvec_device_ms": float(np.median(dev)), "apply_matvec_e2e_ms": float(np.median(wall)),
            "launches": int(launches), "set_operators_ms": set_ms,
            "fiber_force_kernel_ms": float(np.median(k_force)), "fiber_matvec_kernels_ms": float(np.median(k_mv)),
            "fiber_preconditioner_kernel_ms": float(np.median(k_pc)),
            "fiber_preconditioner_GBs": A.nbytes / (float(np.median(k_pc)) * 1e-3) / 1e9,
            "operator_bytes": int(mat_bytes), "host_blas_gemv_ms": cpu_ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak_gbs, "unit": "GB/s",
                         "frac": gbs / hbm_peak_gbs},
            "max_rel_err_fw": e_f, "max_rel_err_res_fibers": e_r}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matvec", action="store_true")
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-fiber-ops", action="store_true")
    ap.add_argument("--no-symmetric", action="store_true", help="plain kernel only (A/B)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    args.warmup = max(args.warmup, 3)

    _log("importing torch")
    import torch
    import torch.distributed as dist
    _log("torch imported")

    import skellysim_b200 as skb
    from skellysim_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_fib, n_shell = workload_sizes(args.workload, world)
    fib, shell, _ = make_suspension(n_fib, n_shell)
    r_src_all = fib
    r_trg_all = np.concatenate([fib, shell])
    n_src, n_trg = r_src_all.shape[0], r_trg_all.shape[0]
    from skellysim_b200.distributed import RankPartition, allgather_strengths, block_range
    part = RankPartition(n_src, n_trg, world, rank)
    src_chunk = part.src_chunk
    s0, s1 = part.src_range
    if part.gathered_rows != n_src:
        raise SystemExit("bench workloads keep n_src divisible by the rank count")
    # Target list of a rank: ALL fiber nodes first (the self-interaction block, evaluated with the symmetric kernel:
    # each rank owns a serpentine set of block rows and produces partial sums for every fiber node), then the rank's
    # block of the remaining (shell) targets.  Partial sums are combined by one all-reduce per step.
    sym_layout = not args.no_symmetric
    if sym_layout:
        n_rem = n_trg - n_src
        rb, re_ = block_range(n_rem, world, rank)
        my_trg = np.ascontiguousarray(np.concatenate([r_src_all, r_trg_all[n_src + rb:n_src + re_]]))
    else:  # plain kernel: every rank owns a block of the whole target list, no reduction needed
        tb, te = part.trg_range
        my_trg = np.ascontiguousarray(r_trg_all[tb:te])
    n_my_trg = my_trg.shape[0]
    rng = np.random.default_rng(7)
    f_all = rng.uniform(-1, 1, (n_src, 3))  # trapezoid-weighted forces, U[-1,1]

    # device state: positions once ("per timestep"); strengths per step
    ctx = skb.Context(1, device_ids=[local_rank])
    if args.no_symmetric:
        ctx.set_symmetric(0)
    ctx.set_sym_partition(rank, world)
    stream = torch.cuda.current_stream().cuda_stream  # kernels are launched on torch's current stream
    d_rsrc = torch.from_numpy(r_src_all).to(dev)
    d_rtrg = torch.from_numpy(my_trg).to(dev)
    # N > 1 with the symmetric layout: the self block (partial sums, needs the reduce-scatter) and the rank's remainder
    # targets live in two contexts, so the remainder runs on a side stream WHILE the fiber rows are reduce-scattered
    two_ctx = sym_layout and world > 1
    ctx_rem, side, ev_g, ev_s = None, None, None, None
    ctx.set_sources_device(skb.KERNEL_STOKESLET, d_rsrc.data_ptr(), n_src, stream)
    if two_ctx:
        ctx.set_targets_device(d_rtrg.data_ptr(), n_src, stream)
        ctx_rem = skb.Context(1, device_ids=[local_rank])
        ctx_rem.set_symmetric(0)
        ctx_rem.set_sources_device(skb.KERNEL_STOKESLET, d_rsrc.data_ptr(), n_src, stream)
        ctx_rem.set_targets_device(d_rtrg.data_ptr() + 24 * n_src, n_my_trg - n_src, stream)
        side = torch.cuda.Stream(device=dev)
        ev_g, ev_s = torch.cuda.Event(), torch.cuda.Event()
    else:
        ctx.set_targets_device(d_rtrg.data_ptr(), n_my_trg, stream)
    d_f_gather = torch.zeros((world * src_chunk, 3), dtype=torch.float64, device=dev)  # all-gather landing zone
    d_f_mine = d_f_gather[rank * src_chunk:(rank + 1) * src_chunk]
    h_f_mine = torch.zeros((src_chunk, 3), dtype=torch.float64).pin_memory()
    h_f_mine[:s1 - s0] = torch.from_numpy(f_all[s0:s1])
    d_f_mine.copy_(h_f_mine, non_blocking=True)
    d_u = torch.empty((max(n_my_trg, 1), 3), dtype=torch.float64, device=dev)
    d_u_fib = d_u[:n_src]
    d_u_mine = torch.empty((src_chunk, 3), dtype=torch.float64, device=dev)  # reduce-scatter output (own fibers)
    h_u = torch.empty((max(n_my_trg, 1), 3), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    # The gathered strength buffer has `world*src_chunk` rows; rows beyond n_src are zero strengths at padded
    # positions only when world*src_chunk > n_src.  Sources are registered with exactly n_src rows, and the
    # gather layout is contiguous by rank, so rows [0, n_src) are the real sources when src_chunk*world == n_src;
    # otherwise the tail ranks are short.  Keep it simple: require divisibility (sizes above are multiples of 32).
    def step_device():
        allgather_strengths(d_f_gather, d_f_mine)  # ONE NCCL all-gather per step (no-op at world == 1)
        if two_ctx:
            ev_g.record()
            side.wait_event(ev_g)
            ctx.eval_device(skb.KERNEL_STOKESLET, d_f_gather.data_ptr(), d_u.data_ptr(), False, stream)
            if n_my_trg > n_src:
                with torch.cuda.stream(side):
                    ctx_rem.eval_device(skb.KERNEL_STOKESLET, d_f_gather.data_ptr(), d_u.data_ptr() + 24 * n_src,
                                        False, side.cuda_stream)
            ev_s.record(side)
            # fiber rows are per-rank partial sums of the symmetric block rows: every rank ends up with the
            # velocities of ITS fibers; the collective overlaps with the remainder targets on the side stream
            dist.reduce_scatter_tensor(d_u_mine, d_u_fib)
            torch.cuda.current_stream().wait_event(ev_s)
        else:
            ctx.eval_device(skb.KERNEL_STOKESLET, d_f_gather.data_ptr(), d_u.data_ptr(), False, stream)

    def step_e2e():
        d_f_mine.copy_(h_f_mine, non_blocking=True)
        step_device()
        h_u.copy_(d_u, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kms = []
        barrier()
        wall0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()  # evict L2 between timed iterations (not timed)
            a.record()
            fn()
            b.record()
            if len(kms) < 8:  # a few kernel-only samples for the roofline (needs the step finished)
                b.synchronize()
                kms.append(ctx.stats()["kernel_ms"])
        barrier()
        wall = time.perf_counter() - wall0
        tot_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, kms

    # correctness gate inside the bench (subset vs the CPU oracle), rank 0
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if sym_layout and not ctx.last_eval_was_symmetric():
        raise SystemExit("symmetric layout requested but the symmetric kernel was not used (problem too small?); "
                         "run with --no-symmetric")
    acc = None
    if rank == 0:
        import oracle as orc
        if world > 1 and sym_layout:  # this rank's final rows: its own fibers (reduce-scatter) + its remainder block
            chk_trg = np.concatenate([r_src_all[s0:s1], my_trg[n_src:]])
            chk_val = np.concatenate([d_u_mine.cpu().numpy()[:s1 - s0], d_u.cpu().numpy()[n_src:]])
        else:
            chk_trg, chk_val = my_trg, d_u.cpu().numpy()
        idx = np.random.default_rng(3).choice(chk_trg.shape[0], size=min(128, chk_trg.shape[0]), replace=False)
        ref = orc.stokeslet_direct_cpu(r_src_all, f_all, chk_trg[idx], 1.0)
        acc = float(np.abs(chk_val[idx] - ref).max() / np.abs(ref).max())

    _log("warm-up + accuracy gate done")
    with ClockSampler(local_rank) as clk:
        n0 = capi.launch_count()
        tot_ms, wall, kms = timed(step_device, args.steps)
        launches = capi.launch_count() - n0
    e2e_path = "pinned host strengths -> H2D -> (all-gather) -> skb_eval_device -> (reduce-scatter) -> D2H velocities"
    if world == 1:
        # the call a SkellySim evaluator makes: skb_eval with HOST buffers (positions cached in a context, strengths
        # in, velocities out; both copies inside the call), pinned buffers, wall clock around the synchronous call
        ctx_host = skb.Context(1, device_ids=[local_rank])
        if args.no_symmetric:
            ctx_host.set_symmetric(0)
        ctx_host.set_targets(my_trg)
        ctx_host.set_sources(skb.KERNEL_STOKESLET, r_src_all)
        f_np = h_f_mine.numpy()[:n_src]
        u_np = h_u.numpy()[:n_my_trg]
        for _ in range(args.warmup):
            ctx_host.eval(skb.KERNEL_STOKESLET, f_np, out=u_np)
        e2e_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx_host.eval(skb.KERNEL_STOKESLET, f_np, out=u_np)
            e2e_ms += 1e3 * (time.perf_counter() - t0)
        if rank == 0:
            ref_sub = d_u.cpu().numpy()
            assert np.abs(u_np - ref_sub).max() <= 1e-12 * np.abs(ref_sub).max(), "host-pointer path disagrees"
        ctx_host.close()
        e2e_path = "skb_eval (C ABI, host pointers, pinned buffers): H2D strengths + kernels + D2H velocities, wall clock"
    else:
        for _ in range(args.warmup):
            step_e2e()
        e2e_ms, _, _ = timed(step_e2e, args.steps)

    pairs_total = float(n_src) * float(n_trg)
    if sym_layout:   # (N > 1: the remainder context runs concurrently inside the symmetric context's kernel interval)
        pairs_rank = float(n_src) * float(n_src) / world + float(n_src) * float(n_my_trg - n_src)
    else:
        pairs_rank = float(n_src) * float(n_my_trg)
    ms_per_step = tot_ms / args.steps
    value = pairs_total / (ms_per_step * 1e-3)
    e2e_val = pairs_total / (e2e_ms / args.steps * 1e-3)
    k_ms = float(np.median(kms)) if kms else float("nan")
    stats = ctx.stats()
    launches_all = launches
    if world > 1:
        t = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        launches_all = int(t.item())

    out = None
    if rank == 0:
        peak = ctx.measure_fp64_peak()
        achieved = SL_FLOP_PER_PAIR * pairs_rank / (k_ms * 1e-3)
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"])
            hbm_src = "measured (MEASURED_PEAKS.json)"
        # algorithmic HBM bytes per launch with cached positions: 24 B/source strengths read
        # + positions 24 B/source + targets 24 B read + 24 B written per target
        alg_bytes = 24.0 * n_src + 24.0 * n_src + 48.0 * n_my_trg
        roofline = {
            "bound": "fp64",
            "kernel": ("pair_sym_kernel<T=4> (self block, 16 FP64 instr/pair) + pair_sum_kernel (remainder)"
                       if sym_layout else f"pair_sum_kernel<stokeslet,T={stats['targets_per_thread']}>"),
            "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "measured on this GPU by skb_measure_fp64_peak (register-resident DFMA loop)",
            "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / 1e12 / NOMINAL_FP64_TFLOPS,
            "flop_per_pair": SL_FLOP_PER_PAIR, "kernel_ms": k_ms,
            "fp64_instr_per_pair": 16 if sym_layout else 22,
            "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": alg_bytes},
            "traffic": None,
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            roofline["traffic"] = json.load(open(prof)).get(args.workload)
        out = {
            "metric": "stokeslet_pair_interactions_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, n_fib, n_shell), "n_src": n_src, "n_trg": n_trg,
                       "pairs_per_step": pairs_total,
                       "parallelism": (f"symmetric block rows (serpentine) + remainder targets partitioned x{world}"
                                       + (", 1 NCCL all-gather of strengths + 1 reduce-scatter of fiber velocities per step"
                                          if world > 1 else "")) if sym_layout else
                       (f"targets+sources block-partitioned x{world}"
                        + (", 1 NCCL all-gather of strengths per step" if world > 1 else "")),
                       "kernel": "Newton's-third-law symmetric kernel on the fiber-fiber block" if sym_layout
                       else "plain kernel",
                       "l2": "flushed between timed steps (256 MiB memset, untimed)",
                       "timing": "per-step CUDA events on the launch stream, summed; max over ranks",
                       "positions": "device-resident across steps (constant within a timestep, system.cpp:486-489)"},
            "e2e": {"value": e2e_val, "unit": "pairs/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(h_f_mine.numel() * 8 * world),
                    "d2h_bytes_per_step": int((world * n_src + (n_trg - n_src)) * 24 if sym_layout else n_trg * 24),
                    "path": e2e_path},
            "gpu_launches": launches_all,
            "launches_per_step": launches_all / args.steps / world,
            "clocks": clk.summary(),
            "roofline": roofline,
            "accuracy": {"max_rel_err_vs_oracle": acc, "targets_checked": 128, "gate": 1e-12},
            "wall_s_timed_region": wall,
        }
        if acc is not None and not (acc < 1e-12):
            out["error"] = f"accuracy gate failed: {acc}"
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg(r_src_all, f_all, r_trg_all)
    _log("headline + e2e + cpu baseline done")
    ctx.close()
    if ctx_rem is not None:
        ctx_rem.close()
    del flush
    mv = None if args.no_matvec else matvec_leg(torch, dist, skb, dev, local_rank, rank, world)
    if rank == 0:
        if mv is not None:
            out["matvec"] = mv
        _log("matvec leg done")
        if world == 1 and not args.no_dense:
            out["periphery_dense"] = dense_leg(skb, out["roofline"]["hbm"]["peak"])
        if world == 1 and not args.no_fiber_ops:
            out["fiber_operators"] = fiber_ops_leg(skb, out["roofline"]["hbm"]["peak"])
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
