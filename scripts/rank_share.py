#!/usr/bin/env python
"""Per-rank device time of an n-way group measured on ONE GPU: member `rank` of `world` evaluates its share of the C3
matvec alone (skb_flow_group_set_solo: no flags, own window only).  usage: rank_share.py [world=8] [rank=0] [workload=c3]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import skellysim_b200 as skb  # noqa: E402
from skellysim_b200 import capi  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    wl = sys.argv[3] if len(sys.argv) > 3 else "c3"
    g = bench.make_system(wl, 1)
    dev = torch.device("cuda", 0)
    n, ns, nb = g["n"], g["shell"].shape[0], g["body"].shape[0]
    f0, f1, s0, s1, b0, b1 = capi.partition_query(g["n_nodes"], ns, nb, world, rank)
    fl = skb.Flow(0)
    fl.set_fibers(g["fib"], g["n_nodes"], g["lengths"])
    fl.set_periphery(g["shell"], g["shell_n"])
    fl.set_bodies(g["body"], g["body_n"], g["centers"])
    fl.set_target_ranges(f0, f1, s0, s1, b0, b1)
    fl.group_init(rank, world)
    fl.group_warmup()
    fl.group_set_solo(True)
    ops = bench.Ops(g, f0, f1)
    fl.set_fiber_class(n, ops.D, ops.P)
    fl.set_fiber_operators(ops.A.base, ops.F.base, ops.xs, ops.lprev, ops.plus, colmajor=True)
    dn = skb.Dense(device_ids=[0])
    dn.set_matrix(skb.DENSE_STRESSLET_PLUS_COMPLEMENTARY, bench.dense_rows(3 * (s1 - s0), 3 * ns, 3 * s0))
    inp = bench.make_inputs(g, 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = dict(x=t(inp["x"][4 * f0 * n:4 * f1 * n]), xs=t(inp["xs"][s0:s1]), bd=t(inp["bd"]), f=t(inp["ft"][:, :3]),
             t=t(inp["ft"][:, 3:]), link=t(inp["link"][f0:f1]))
    res = torch.zeros(4 * (f1 - f0) * n, dtype=torch.float64, device=dev)
    outs = torch.zeros((s1 - s0, 3), dtype=torch.float64, device=dev)
    vb = torch.zeros((max(b1 - b0, 1), 3), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def mv():
        fl.apply_matvec_device(dn, d["x"].data_ptr(), d["xs"].data_ptr(), d["bd"].data_ptr(), d["f"].data_ptr(),
                               d["t"].data_ptr(), d["link"].data_ptr(), 1.0, res.data_ptr(), outs.data_ptr(),
                               vb.data_ptr(), st)
    for _ in range(5):
        mv()
    torch.cuda.synchronize()
    reps = int(os.environ.get("REPS", "40"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        mv()
    e1.record()
    torch.cuda.synchronize()
    sym_ms, sym_pairs = fl.last_sym_kernel()
    print(json.dumps({"world": world, "rank": rank, "workload": wl, "ms_per_matvec_share": e0.elapsed_time(e1) / reps,
                      "sym_kernel_ms": sym_ms, "sym_pairs_per_s": sym_pairs / (sym_ms * 1e-3) if sym_ms else None,
                      "launches": fl.stats()["launches"], "ideal_ms": None}))


if __name__ == "__main__":
    main()
