#!/usr/bin/env python
"""Does the background row streamer share the SMs with the symmetric pair kernel?  (GPU box, one device.)
Times, with CUDA events on each stream: the fiber-fiber flow alone (stream A), the dense product alone (stream B, streamer
and classic kernel), and both launched together.  Perfect overlap: t_both(A) ~ t_alone(A) and t_both(B) <= t_both(A).
usage: probe_overlap.py [n_fibers=3000] [n_shell_rows=18000]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skellysim_b200 as skb  # noqa: E402


def main():
    nfib = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 18000
    rng = np.random.default_rng(0)
    n = 32
    pos = rng.uniform(-1, 1, (nfib * n, 3))
    dev = torch.device("cuda", 0)
    fl = skb.Flow(0)
    fl.set_fibers(pos, [n] * nfib, np.ones(nfib))
    fl.set_periphery(np.zeros((0, 3)), np.zeros((0, 3)))
    fl.set_bodies(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)))
    f = torch.from_numpy(rng.uniform(-1, 1, (nfib * n, 3))).to(dev)
    v = torch.zeros_like(f)
    dn = skb.Dense(device_ids=[0])
    A = torch.randn(rows, rows, dtype=torch.float64).numpy()
    dn.set_matrix(skb.DENSE_M_INV, A)
    x = torch.randn(rows, dtype=torch.float64, device=dev)
    y = torch.zeros(rows, dtype=torch.float64, device=dev)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)

    def flow():
        fl.matvec_device(f.data_ptr(), 0, 0, 0, 0, 1.0, v.data_ptr(), sa.cuda_stream)

    def stream():
        dn.apply_background_device(skb.DENSE_M_INV, x.data_ptr(), y.data_ptr(), sb.cuda_stream)

    def classic():
        dn.apply_device(skb.DENSE_M_INV, x.data_ptr(), 0, y.data_ptr(), sb.cuda_stream)

    def timed(fa, fb, reps=5):
        ta, tb = [], []
        for _ in range(reps):
            torch.cuda.synchronize()
            ea0, ea1, eb0, eb1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            if fb:
                eb0.record(sb)
                fb()
                eb1.record(sb)
            if fa:
                ea0.record(sa)
                fa()
                ea1.record(sa)
            torch.cuda.synchronize()
            if fa:
                ta.append(ea0.elapsed_time(ea1))
            if fb:
                tb.append(eb0.elapsed_time(eb1))
        med = lambda t: round(float(np.median(t)), 4) if t else None
        return med(ta), med(tb)

    for _ in range(3):
        flow(); stream(); classic()
    torch.cuda.synchronize()
    out = {"lib": os.environ.get("SKB_LIBRARY", "default"), "n_nodes": nfib * n, "rows": rows,
           "gbytes": rows * rows * 8 / 1e9}
    out["flow_alone_ms"] = timed(flow, None)[0]
    out["streamer_alone_ms"] = timed(None, stream)[1]
    out["classic_alone_ms"] = timed(None, classic)[1]
    out["both_streamer_ms"] = timed(flow, stream)
    out["both_classic_ms"] = timed(flow, classic)
    out["sym_kernel_ms_last"] = fl.last_sym_kernel()[0]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
