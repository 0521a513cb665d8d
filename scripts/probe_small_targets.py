#!/usr/bin/env python
"""Tuning probe (GPU box): the plain kernel with FEW targets against many sources (a rank's remainder rows under 8-way
sharding: 800 x 96 000; the periphery's stresslet on a rank's fibers: 12 050 x 6 000) over (T, splits)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skellysim_b200 as skb  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    out = []
    with skb.Context(1) as ctx:
        ctx.set_symmetric(0)
        for kind, fdim, ns, nt in ((0, 3, 96000, 800), (0, 3, 96000, 1600), (1, 9, 6000, 12050), (1, 9, 6000, 24100)):
            rs, rt = rng.uniform(-1, 1, (ns, 3)), rng.uniform(-1, 1, (nt, 3))
            f = rng.uniform(-1, 1, (ns, fdim))
            ctx.set_targets(rt)
            ctx.set_sources(kind, rs)
            for T in (0, 1, 2, 4, 8):
                for S in ((0,) if T == 0 else (0, 32, 64, 100, 150, 200, 250) if kind == 0 else (0, 2, 4, 6, 8, 12, 16, 24)):
                    ctx.set_tuning(T, S)
                    ctx.eval(kind, f)
                    ks = []
                    for _ in range(5):
                        ctx.eval(kind, f)
                        ks.append(ctx.stats()["kernel_ms"])
                    st = ctx.stats()
                    rec = dict(kind=kind, n_src=ns, n_trg=nt, T=st["targets_per_thread"], S=st["source_splits"],
                               ctas=st["grid_ctas"], forced=(T, S), us=round(1e3 * float(np.median(ks)), 1),
                               gpairs=round(ns * nt / float(np.median(ks)) / 1e6, 1))
                    out.append(rec)
                    print(json.dumps(rec), flush=True)
            ctx.set_tuning(0, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_small_targets.json", "w"))


if __name__ == "__main__":
    main()
