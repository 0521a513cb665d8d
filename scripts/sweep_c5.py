#!/usr/bin/env python
"""BASELINE C5 / SURVEY.md 8d S2: Stokeslet throughput sweep n_src = n_trg in {1e3 .. 1e6} (targets == sources, so the
self pairs are exercised and the symmetric kernel engages from 4096 nodes) plus the rectangular 1e4 x 1e6 and
1e6 x 1e4 shapes, on one GPU, with the accuracy gate and the CPU port of kernels::stokeslet_direct_cpu timed in the
same run (bounded target sample).  Measurement script (uses oracle/ as checker and CPU baseline)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
import skellysim_b200 as skb  # noqa: E402


def cpu_rate(rs, f, rt, budget_s=1.5):
    n = min(rt.shape[0], 256)
    orc.stokeslet_direct_cpu(rs, f, rt[:n], 1.0)
    t0 = time.perf_counter()
    orc.stokeslet_direct_cpu(rs, f, rt[:n], 1.0)
    dt = time.perf_counter() - t0
    n2 = int(min(rt.shape[0], max(n, n * budget_s / max(dt, 1e-6))))
    t0 = time.perf_counter()
    orc.stokeslet_direct_cpu(rs, f, rt[:n2], 1.0)
    dt = time.perf_counter() - t0
    return rs.shape[0] * n2 / dt, n2


def main():
    shapes = [(n, n) for n in (1000, 3000, 10000, 30000, 100000, 300000, 1000000)] + [(10000, 1000000),
                                                                                       (1000000, 10000)]
    out = []
    with skb.Context(1) as c:
        for seed, (ns, nt) in enumerate(shapes, 1):
            rng = np.random.default_rng(seed % 3 + 1)
            rs = rng.uniform(-1, 1, (ns, 3))
            rt = rs if ns == nt else rng.uniform(-1, 1, (nt, 3))
            f = rng.uniform(-1, 1, (ns, 3))
            c.set_targets(rt)
            c.set_sources(0, rs)
            u = c.eval(0, f)
            reps = 3 if ns * nt >= 1e11 else 7
            ks, ws = [], []
            for _ in range(reps):
                t0 = time.perf_counter()
                u = c.eval(0, f)
                ws.append(1e3 * (time.perf_counter() - t0))
                ks.append(c.stats()["kernel_ms"])
            sub = np.random.default_rng(0).choice(nt, min(64, nt), replace=False)
            ref = orc.stokeslet_direct_cpu(rs, f, rt[sub], 1.0)
            err = float(np.abs(u[sub] - ref).max() / np.abs(ref).max())
            cpu, n_cpu = cpu_rate(rs, f, rt)
            k, w = float(np.median(ks)), float(np.median(ws))
            rec = dict(n_src=ns, n_trg=nt, kernel_ms=k, call_ms=w, gpairs_s_kernel=ns * nt / k / 1e6,
                       gpairs_s_call=ns * nt / w / 1e6, symmetric=c.last_eval_was_symmetric(),
                       max_rel_err=err, cpu_port_gpairs_s=cpu / 1e9, cpu_sample_targets=n_cpu,
                       cpu_threads=orc.max_threads())
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sweep_c5.json", "w"), indent=1)


if __name__ == "__main__":
    main()
