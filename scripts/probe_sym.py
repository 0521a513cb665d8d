#!/usr/bin/env python
"""Tuning probe (GPU box): kernel-only time of the symmetric Stokeslet self-interaction for the library named by
SKB_LIBRARY (one build variant per process), with an accuracy check against the CPU oracle on a target subset.
usage: probe_sym.py [n ...]   (default 96000 32000)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
import skellysim_b200 as skb  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [96000, 32000]
    rng = np.random.default_rng(0)
    out = {"lib": os.environ.get("SKB_LIBRARY", "default"), "results": []}
    with skb.Context(1) as ctx:
        ctx.set_symmetric(1)
        for n in sizes:
            r = rng.uniform(-1, 1, (n, 3))
            f = rng.uniform(-1, 1, (n, 3))
            ctx.set_targets(r)
            ctx.set_sources(skb.KERNEL_STOKESLET, r)
            u = ctx.eval(skb.KERNEL_STOKESLET, f)
            assert ctx.last_eval_was_symmetric()
            ks = []
            for _ in range(7):
                ctx.eval(skb.KERNEL_STOKESLET, f)
                ks.append(ctx.stats()["kernel_ms"])
            k = float(np.median(ks))
            idx = rng.choice(n, 96, replace=False)
            ref = orc.stokeslet_direct_cpu(r, f, r[idx], 1.0)
            err = float(np.abs(u[idx] - ref).max() / np.abs(ref).max())
            out["results"].append({"n": n, "kernel_ms": round(k, 4), "min_ms": round(min(ks), 4),
                                   "gpairs_s": round(n * n / k / 1e6, 1), "err": err})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
