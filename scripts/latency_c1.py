#!/usr/bin/env python
"""BASELINE C1 (16 fibers x 32 nodes = 512 nodes, targets == sources): call latency of the drop-in paths vs the CPU
port of the reference's OpenMP direct path (oracle/; measurement script)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
import skellysim_b200 as skb  # noqa: E402


def wall(fn, reps=300):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e6 * (time.perf_counter() - t0) / reps


rng = np.random.default_rng(1)
out = {}
for n_fib in (16, 64, 256):
    n = n_fib * 32
    pos = []
    for _ in range(n_fib):
        x0 = rng.uniform(-2, 2, 3)
        nh = rng.normal(size=3)
        nh /= np.linalg.norm(nh)
        pos.append(x0 + np.linspace(0, 1, 32)[:, None] * nh)
    r = np.concatenate(pos)
    f = rng.uniform(-1, 1, (n, 3))
    e3 = np.zeros((0, 3))
    rec = {"n_nodes": n}
    with skb.Context(1) as c:
        c.set_targets(r)
        c.set_sources(0, r)
        rec["skb_eval_us"] = wall(lambda: c.eval(0, f))
        rec["skb_eval_device_ms_stat"] = c.stats()["total_ms"] * 1e3
    rec["skb_stokeslet_direct_us"] = wall(lambda: skb.stokeslet_direct(r, f, r))
    with skb.Flow(0) as fl:
        fl.set_fibers(r, [32] * n_fib, [1.0] * n_fib)
        fl.set_periphery(e3, e3)
        fl.set_bodies(e3, e3, e3)
        ft = np.zeros((0, 6))
        rec["skb_flow_matvec_us"] = wall(lambda: fl.matvec(f, e3, e3, ft, 1.0))
        t6 = rng.uniform(-2, 2, (6, 3))
        rec["skb_flow_velocity_at_6_targets_us"] = wall(lambda: fl.velocity_at_targets(t6, f, e3, e3, ft, 1.0))
    rec["cpu_port_all_threads_us"] = wall(lambda: orc.stokeslet_direct_cpu(r, f, r, 1.0), reps=100)
    rec["cpu_port_1_thread_us"] = wall(lambda: orc.stokeslet_direct_cpu(r, f, r, 1.0, 1), reps=20)
    if orc.refgpu_available():
        rec["reference_kernels_cu_us"] = wall(lambda: orc.ref_stokeslet_direct_gpu_impl(r, f, r), reps=100)
    print(json.dumps(rec), flush=True)
    out[n] = rec
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/latency_c1.json", "w"), indent=1)
