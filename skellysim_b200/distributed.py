"""One-rank-per-GPU sharding of the pair-kernel path (used by bench.py; CPU-testable with gloo).

Targets are independent (each u_t is a private sum over all sources; the reference already exploits this across
OpenMP threads, kernels.cpp:58-65), so targets AND sources are block-partitioned over ranks and the only exchange
is ONE all-gather of the source strengths per evaluation.  Positions are all-gathered once per
timestep.  This module only moves strengths between ranks; it computes no velocities."""
from __future__ import annotations

from dataclasses import dataclass


def block_range(n: int, world: int, rank: int):
    """Contiguous block [b, e) of rank `rank` when n items are split into `world` blocks of ceil(n/world)."""
    chunk = -(-n // world) if world > 0 else n
    b = min(n, rank * chunk)
    e = min(n, (rank + 1) * chunk)
    return b, e


@dataclass
class RankPartition:
    n_src: int
    n_trg: int
    world: int
    rank: int

    @property
    def src_chunk(self) -> int:  # all-gather slot size (equal on every rank)
        return -(-self.n_src // self.world)

    @property
    def src_range(self):
        return block_range(self.n_src, self.world, self.rank)

    @property
    def trg_range(self):
        return block_range(self.n_trg, self.world, self.rank)

    @property
    def gathered_rows(self) -> int:
        return self.src_chunk * self.world


def allgather_strengths(gathered, mine, group=None):
    """In-place all-gather of this rank's strength slot into `gathered` ((world*src_chunk, d) tensor; `mine` must be
    the view gathered[rank*src_chunk:(rank+1)*src_chunk]).  NCCL on GPUs, gloo on CPU."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_gather_into_tensor(gathered, mine, group=group)
    return gathered


def sym_row_owner(block_row: int, n_parts: int) -> int:
    """Owner of block row I of the symmetric kernel's upper triangle (mirror of sym_row_owner in
    csrc/pair_kernels.cuh): serpentine over rows, so every rank gets the same mix of long and short rows."""
    k, m = divmod(block_row, n_parts)
    return n_parts - 1 - m if (k & 1) else m


def sym_block_pairs(n_blocks: int, part: int, n_parts: int):
    """(I, J) block pairs a rank evaluates in both directions: rows it owns of the strict upper triangle, plus its
    diagonal blocks (I, I)."""
    for i in range(n_blocks):
        if sym_row_owner(i, n_parts) != part:
            continue
        yield i, i
        for j in range(i + 1, n_blocks):
            yield i, j


def reference_counts(n: int, world: int):
    """Block sizes of the reference's MPI decompositions: n // world each, the first n % world ranks one more
    (fibers: fiber_container_finite_difference.cpp:102-108; periphery nodes: periphery.cpp:387-400)."""
    return [n // world + (1 if r < n % world else 0) for r in range(world)]


def reference_rank_ranges(n_fibers: int, n_shell_nodes: int, n_body_nodes: int, rank: int, world: int):
    """The rows of [fibers | shell | bodies] one MPI rank of the reference owns, as the six arguments of
    skb_flow_set_target_ranges: whole fibers [f0, f1), periphery nodes [s0, s1), and all body nodes on rank 0
    (body_container.cpp keeps the bodies on world_rank_ 0)."""
    def block(n):
        c = reference_counts(n, world)
        b = sum(c[:rank])
        return b, b + c[rank]
    f0, f1 = block(n_fibers)
    s0, s1 = block(n_shell_nodes)
    b0, b1 = (0, n_body_nodes) if rank == 0 else (0, 0)
    return f0, f1, s0, s1, b0, b1


def allgatherv_rows(local, counts, group=None):
    """All-gather of per-rank row blocks of unequal length (fw of the own fibers -> fw of all fibers): pads every
    block to max(counts) rows, one all_gather_into_tensor, returns the concatenation without the padding.
    `local` is a (counts[rank], d) tensor; NCCL on GPUs, gloo on CPU."""
    import torch
    import torch.distributed as dist
    world = len(counts)
    if world == 1 or not dist.is_initialized():
        return local
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)])
