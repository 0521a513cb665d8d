#!/usr/bin/env python
"""GPU box: wall-clock of the UNMODIFIED reference CUDA path (oracle/_ref/libskelly_ref_kernels_cu.so ==
src/core/kernels.cu: malloc + H2D + tiled_driver + D2H + free per call, kernels.cu:148-178) against the drop-in's
stateless entry point with the same semantics, same host buffers.  Test/measurement script: uses oracle/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
import skellysim_b200 as skb  # noqa: E402

shapes = [(512, 512), (1229, 743), (10000, 10000), (32000, 40000), (96000, 102400)]
rng = np.random.default_rng(0)
out = []
for ns, nt in shapes:
    rs, rt, f3, f9 = rng.uniform(-1, 1, (ns, 3)), rng.uniform(-1, 1, (nt, 3)), rng.uniform(-1, 1, (ns, 3)), \
        rng.uniform(-1, 1, (ns, 9))
    rec = dict(n_src=ns, n_trg=nt)
    for name, fn, f in (("ref_sl", orc.ref_stokeslet_direct_gpu_impl, f3), ("new_sl", skb.stokeslet_direct, f3),
                        ("ref_dl", orc.ref_stresslet_direct_gpu_impl, f9), ("new_dl", skb.stresslet_direct, f9)):
        fn(rs, f, rt)
        reps = 3 if ns * nt > 1e9 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            u = fn(rs, f, rt)
        rec[name + "_ms"] = 1e3 * (time.perf_counter() - t0) / reps
    rec["speedup_sl"] = rec["ref_sl_ms"] / rec["new_sl_ms"]
    rec["speedup_dl"] = rec["ref_dl_ms"] / rec["new_dl_ms"]
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ref_gpu_timing.json", "w"), indent=1)
