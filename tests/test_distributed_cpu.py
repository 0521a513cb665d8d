"""world_size-2 gloo tests (CPU) of the N>1 sharding used by bench.py: block partition of sources and targets, one
all-gather of strengths, per-rank target blocks that tile the global result.  The pair sums themselves are evaluated
with the CPU oracle here (this is a test of the sharding logic, not of the CUDA kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_src, n_trg, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as orc
    from skellysim_b200.distributed import RankPartition, allgather_strengths
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # identical global problem on every rank
    r_src, r_trg = rng.uniform(-1, 1, (n_src, 3)), rng.uniform(-1, 1, (n_trg, 3))
    f = rng.uniform(-1, 1, (n_src, 3))
    part = RankPartition(n_src, n_trg, world, rank)
    gathered = torch.zeros((part.gathered_rows, 3), dtype=torch.float64)
    mine = gathered[rank * part.src_chunk:(rank + 1) * part.src_chunk]
    b, e = part.src_range
    mine[:e - b] = torch.from_numpy(f[b:e])          # each rank only knows its own strengths
    allgather_strengths(gathered, mine)
    f_all = gathered.numpy()[:n_src]
    assert np.array_equal(f_all, f), "all-gather layout must reproduce the global strength array"
    tb, te = part.trg_range
    u_block = orc.stokeslet_direct(r_src, f_all, r_trg[tb:te])
    np.save(os.path.join(out_dir, f"u_{rank}.npy"), u_block)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_src,n_trg", [(64, 50), (33, 7), (5, 1)])
def test_two_rank_sharding_tiles_the_global_result(tmp_path, n_src, n_trg):
    import oracle as orc
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_src, n_trg, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    r_src, r_trg = rng.uniform(-1, 1, (n_src, 3)), rng.uniform(-1, 1, (n_trg, 3))
    f = rng.uniform(-1, 1, (n_src, 3))
    ref = orc.stokeslet_direct(r_src, f, r_trg)
    got = np.concatenate([np.load(tmp_path / f"u_{r}.npy") for r in range(world)])
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_block_range_covers_everything():
    from skellysim_b200.distributed import block_range
    for n in (0, 1, 7, 8, 9, 1000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = block_range(n, world, r)
                assert 0 <= b <= e <= n
                cover += list(range(b, e))
            assert cover == list(range(n))


def _sym_worker(rank, world, port, n, block, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as orc
    from skellysim_b200.distributed import sym_block_pairs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    r = rng.uniform(-1, 1, (n, 3))
    f = rng.uniform(-1, 1, (n, 3))
    nb = -(-n // block)
    part = np.zeros((n, 3))
    # each rank: its block rows, both directions of every (I, J) pair from ONE geometry pass (here: two oracle calls)
    for i, j in sym_block_pairs(nb, rank, world):
        si, sj = slice(i * block, min(n, (i + 1) * block)), slice(j * block, min(n, (j + 1) * block))
        part[si] += orc.stokeslet_direct(r[sj], f[sj], r[si])        # forward: u_I += G f_J
        if i != j:
            part[sj] += orc.stokeslet_direct(r[si], f[si], r[sj])    # reverse: u_J += G f_I
    t = torch.from_numpy(part)
    dist.all_reduce(t)                                                # the one reduction of the symmetric layout
    if rank == 0:
        np.save(os.path.join(out_dir, "u_sym.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_symmetric_row_partition_reduces_to_the_full_self_interaction(tmp_path, world):
    import oracle as orc
    n, block = 300, 32
    mp.spawn(_sym_worker, args=(world, _free_port(), n, block, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(9)
    r = rng.uniform(-1, 1, (n, 3))
    f = rng.uniform(-1, 1, (n, 3))
    ref = orc.stokeslet_direct(r, f, r)
    got = np.load(tmp_path / "u_sym.npy")
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-13


def test_sym_row_owner_is_balanced_and_complete():
    from skellysim_b200.distributed import sym_block_pairs, sym_row_owner
    for nb in (1, 2, 7, 64, 177):
        for parts in (1, 2, 4, 8):
            owners = [sym_row_owner(i, parts) for i in range(nb)]
            assert all(0 <= o < parts for o in owners)
            seen = set()
            load = [0] * parts
            for p in range(parts):
                for i, j in sym_block_pairs(nb, p, parts):
                    assert (i, j) not in seen and i <= j
                    seen.add((i, j))
                    load[p] += 1
            assert len(seen) == nb * (nb + 1) // 2          # every unordered block pair exactly once
            if nb >= 8 * parts:
                assert max(load) - min(load) <= 0.05 * max(load) + parts
