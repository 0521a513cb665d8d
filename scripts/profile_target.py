#!/usr/bin/env python
"""Small fixed workload for ncu captures: C3-shaped SL and DL evaluations.
usage: profile_target.py [reps] [T] [S] [kinds]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skellysim_b200 as skb  # noqa: E402

ns, nt = 96000, 102400
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
S = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kinds = sys.argv[4] if len(sys.argv) > 4 else "sd"
rng = np.random.default_rng(0)
rs = rng.uniform(-1, 1, (ns, 3))
rt = np.concatenate([rs, rng.uniform(-1, 1, (nt - ns, 3))])
with skb.Context(1) as ctx:
    ctx.set_tuning(T, S)
    ctx.set_targets(rt)
    ctx.set_sources(0, rs)
    ctx.set_sources(1, rs)
    f3, f9 = rng.uniform(-1, 1, (ns, 3)), rng.uniform(-1, 1, (ns, 9))
    for _ in range(reps):
        if "s" in kinds:
            ctx.eval(0, f3)
            print(ctx.stats())
        if "d" in kinds:
            ctx.eval(1, f9)
            print(ctx.stats())
