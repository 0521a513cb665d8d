#!/usr/bin/env python
"""Strong-scaling timing of the FULL System::apply_matvec (system.cpp:269-324) with one rank per GPU in the reference's
own decomposition: whole fibers / periphery-node blocks / bodies on rank 0 (skb_flow_set_target_ranges), own-fiber
operators resident per rank, all-gather of fw and of the shell density per matvec (RankApplyMatvec).

    python scripts/bench_rank_apply_matvec.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_rank_apply_matvec.py --steps 20

STATUS: the sequencing and the collectives are covered on the CPU with gloo (tests/test_distributed_cpu.py) and every
device call is covered rank by rank on one GPU (tests/test_gpu_fiberops.py); this script itself has not been timed on
several GPUs yet (round-1 GPU budget) -- its numbers are not quoted anywhere.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import skellysim_b200 as skb
    from bench import make_c3_system
    from skellysim_b200.distributed import RankApplyMatvec

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = make_c3_system()
    n, n_fibers, eta = 32, 3000, 1.0
    nf, ns, nb = g["fib"].shape[0], g["shell"].shape[0], g["body"].shape[0]
    fl = skb.Flow(local_rank)
    fl.set_fibers(g["fib"], g["n_nodes"], g["lengths"])
    fl.set_periphery(g["shell"], g["shell_n"])
    fl.set_bodies(g["body"], g["body_n"], g["centers"])
    ram = RankApplyMatvec(fl, g["n_nodes"], ns, nb, 1, rank, world, device=dev)
    f0, f1, s0, s1, b0, b1 = ram.ranges
    fl.set_target_ranges(*ram.ranges)
    rng = np.random.default_rng(100 + rank)          # own operators only
    k = f1 - f0
    A = rng.standard_normal((k, 4 * n, 4 * n)) / np.sqrt(4 * n)
    F = rng.standard_normal((k, 3 * n, 4 * n)) / np.sqrt(4 * n)
    xs = rng.standard_normal((k * n, 3))
    xs /= np.linalg.norm(xs, axis=1)[:, None]
    grng = np.random.default_rng(7)                   # class matrices: identical on every rank
    fl.set_fiber_class(n, grng.standard_normal((n, n)), grng.standard_normal((4 * n - 14, 4 * n)) / np.sqrt(4 * n))
    fl.set_fiber_operators(A, F, xs, np.ones(k), rng.integers(0, 2, k).astype(np.int32))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    x_own = t(rng.standard_normal(4 * k * n))
    xs_own = t(g["sd"][s0:s1])
    z = lambda a: t(a) if rank == 0 else torch.zeros(a.shape, dtype=torch.float64, device=dev)
    bd, bf, bt = z(g["bd"]), z(g["force"]), z(g["torque"])
    link = t(rng.standard_normal((k, 7)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        ram.apply(x_own, xs_own, bd, bf, bt, link, eta)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    for a, b in evs:
        flush.zero_()
        a.record()
        res, v_s, v_b = ram.apply(x_own, xs_own, bd, bf, bt, link, eta)
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ok = bool(torch.isfinite(res).all() and torch.isfinite(v_s).all())
    if rank == 0:
        print(json.dumps({"metric": "apply_matvec_ms", "value": float(tt.item()), "unit": "ms", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "higher_is_better": False, "scaling": "strong",
                          "finite": ok,
                          "config": {"workload": "c3: 3000 fibers x 32 + 6000 periphery nodes + 1 body x 400 nodes; "
                                                 "full apply_matvec per rank (own-fiber operators, all-gather of fw "
                                                 "and shell density, flow matvec over own rows, fc.matvec)"}}))
    fl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
