#!/usr/bin/env python
"""The two per-fiber operator launches of the C3 matvec alone (GPU box): CUDA-event times and HBM fractions.
usage: probe_fiber_ops.py [reps=20]   (run it under ncu -k regex:fiber_gemv to look inside the kernels)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import skellysim_b200 as skb  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    g = bench.make_system("c3", 1)
    dev = torch.device("cuda", 0)
    n = g["n"]
    fl = skb.Flow(0)
    fl.set_fibers(g["fib"], g["n_nodes"], g["lengths"])
    fl.set_periphery(np.zeros((0, 3)), np.zeros((0, 3)))
    fl.set_bodies(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)))
    ops = bench.Ops(g, 0, g["n_fibers"])
    fl.set_fiber_class(n, ops.D, ops.P)
    fl.set_fiber_operators(ops.A.base, ops.F.base, ops.xs, ops.lprev, ops.plus, colmajor=True)
    nf = g["fib"].shape[0]
    x = torch.randn(4 * nf, dtype=torch.float64, device=dev)
    v = torch.randn(nf, 3, dtype=torch.float64, device=dev)
    link = torch.randn(g["n_fibers"], 7, dtype=torch.float64, device=dev)
    fw = torch.zeros(nf, 3, dtype=torch.float64, device=dev)
    res = torch.zeros(4 * nf, dtype=torch.float64, device=dev)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    kf, km = [], []
    for _ in range(reps + 2):
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        fl.apply_fiber_force_device(x.data_ptr(), fw.data_ptr(), st)
        b.record()
        fl.fiber_matvec_device(x.data_ptr(), v.data_ptr(), link.data_ptr(), res.data_ptr(), st)
        c.record()
        torch.cuda.synchronize()
        kf.append(a.elapsed_time(b))
        km.append(b.elapsed_time(c))
    kf, km = float(np.median(kf[2:])), float(np.median(km[2:]))
    print(json.dumps({"lib": os.environ.get("SKB_LIBRARY", "default"), "force_ms": kf, "matvec_ms": km,
                      "force_gbs": ops.F.nbytes / kf / 1e6, "matvec_gbs": ops.A.nbytes / km / 1e6,
                      "leg_frac_of_6477": (ops.A.nbytes + ops.F.nbytes) / (kf + km) / 1e6 / 6477.4}))


if __name__ == "__main__":
    main()
