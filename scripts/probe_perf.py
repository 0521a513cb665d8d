#!/usr/bin/env python
"""Tuning probe (GPU box): kernel-only time of the pair kernels for each (T, S) variant."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skellysim_b200 as skb  # noqa: E402


def run(ctx, kind, f, reps=5):
    ctx.eval(kind, f)
    ks = []
    for _ in range(reps):
        ctx.eval(kind, f)
        ks.append(ctx.stats())
    k = sorted(s["kernel_ms"] for s in ks)[len(ks) // 2]
    t = sorted(s["total_ms"] for s in ks)[len(ks) // 2]
    return k, t, ks[-1]


QUICK = "--quick" in sys.argv
SPLITS = (0, 4, 8, 16)
for _a in sys.argv:
    if _a.startswith("--splits="):
        SPLITS = tuple(int(x) for x in _a.split("=")[1].split(","))


def main():
    shapes = [(32000, 40000), (96000, 102400)]
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if args:
        shapes = [tuple(int(x) for x in a.split("x")) for a in args]
    rng = np.random.default_rng(0)
    out = []
    with skb.Context(1) as ctx:
        peak = ctx.measure_fp64_peak()
        print(f"fp64 DFMA probe: {peak/1e12:.2f} TFLOP/s", flush=True)
        for ns, nt in shapes:
            rs = rng.uniform(-1, 1, (ns, 3))
            rt = np.concatenate([rs, rng.uniform(-1, 1, (nt - ns, 3))]) if nt >= ns else rng.uniform(-1, 1, (nt, 3))
            ctx.set_targets(rt)
            for kind, fdim, flop, name in (((0, 3, 28, "SL"),) if "--sl-only" in sys.argv else ((0, 3, 28, "SL"), (1, 9, 40, "DL"))):
                ctx.set_sources(kind, rs)
                f = rng.uniform(-1, 1, (ns, fdim))
                for T in (-1, 0, 1, 2, 4, 8):
                    for S in ((0,) if (T <= 0 or QUICK) else SPLITS):
                        if T == -1:  # symmetric (Newton's third law) kernel, Stokeslet self-interaction only
                            if kind != 0 or nt < ns:
                                continue
                            ctx.set_symmetric(1)
                            ctx.set_tuning(0, 0)
                        else:
                            ctx.set_symmetric(0)
                            ctx.set_tuning(T, S)
                        k, t, st = run(ctx, kind, f)
                        pairs = ns * nt
                        rec = dict(kind=name + ("_sym" if T == -1 else ""), n_src=ns, n_trg=nt, T=st["targets_per_thread"], S=st["source_splits"],
                                   ctas=st["grid_ctas"], forced=(T, S), kernel_ms=round(k, 4), total_ms=round(t, 4),
                                   gpairs_s=round(pairs / k / 1e6, 1), tflops=round(flop * pairs / k / 1e9, 2),
                                   frac_probe=round(flop * pairs / (k * 1e-3) / peak, 3))
                        out.append(rec)
                        print(json.dumps(rec), flush=True)
                ctx.set_tuning(0, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(fp64_peak=peak, results=out), open("gpurun_out/probe_perf.json", "w"), indent=1)


if __name__ == "__main__":
    main()
