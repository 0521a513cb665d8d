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


def connect_group(flow, rank: int, world: int, group=None):
    """One rank per GPU: make `flow` (geometry and target ranges already set) a member of a peer-memory group.  Every
    rank exports the CUDA-IPC handle of its window, the 64-byte handles travel through torch.distributed once, every
    rank maps all peers' windows.  After this, flow.matvec_device / apply_matvec_device exchange strengths and partial
    velocities through NVLink loads / stores inside the library's own kernels -- no collective call per matvec."""
    import torch.distributed as dist
    flow.group_init(rank, world)
    if world == 1:
        return
    handles = [None] * world
    dist.all_gather_object(handles, flow.group_export(), group=group)
    for r in range(world):
        if r != rank:
            flow.group_import(r, handles[r])
    flow.group_warmup()  # every buffer at its final size before the first flag wait
    dist.barrier(group=group)


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


class RankApplyMatvec:
    """One rank of System::apply_matvec (system.cpp:269-324) in the reference's decomposition, one rank per GPU.

    The rank owns whole fibers, a block of periphery nodes and (rank 0) the bodies (`reference_rank_ranges`); it holds
    only its own part of the solution vector, like the reference's local vectors.  Per matvec:

        fw_own   = force_operator_ * x_own                         (device, own fibers)
        fw_all   = all-gather(fw_own);  x_shell_all = all-gather(x_shell_own);  body strengths broadcast from rank 0
        v_own    = flow matvec over [own fiber nodes | own shell rows | own body rows]   (device)
        res_own  = A_ x_own - P vT(v_own) + xs_vT + y_BC            (device, own fibers)

    `flow` is a capi.Flow (or anything with its *_device methods) that already holds the FULL geometry and, after
    `set_target_ranges(*self.ranges)`, the operators of the own fibers.  All arrays are torch tensors on the flow's
    device; nothing is copied to the host.  This class only moves strengths between ranks and sequences the calls.
    """

    def __init__(self, flow, fiber_n_nodes, n_shell_nodes, n_body_nodes, n_bodies, rank, world, device="cpu",
                 group=None):
        import torch
        self.flow, self.rank, self.world, self.group = flow, rank, world, group
        self.torch = torch
        off = [0]
        for n in fiber_n_nodes:
            off.append(off[-1] + int(n))
        self.ranges_all = [reference_rank_ranges(len(fiber_n_nodes), n_shell_nodes, n_body_nodes, r, world)
                           for r in range(world)]
        self.ranges = self.ranges_all[rank]
        self.fib_counts = [off[r[1]] - off[r[0]] for r in self.ranges_all]       # fiber nodes per rank
        self.shell_counts = [r[3] - r[2] for r in self.ranges_all]
        self.n_fib, self.n_shell, self.n_body, self.n_bodies = off[-1], n_shell_nodes, n_body_nodes, n_bodies
        f0, f1, s0, s1, b0, b1 = self.ranges
        self.n_fib_own, self.n_shell_own, self.n_body_own = self.fib_counts[rank], s1 - s0, b1 - b0
        self.n_fibers_own = f1 - f0
        kw = dict(dtype=torch.float64, device=device)
        self._fw_own = torch.zeros((max(self.n_fib_own, 1), 3), **kw)
        self._fw_all = torch.zeros((max(self.n_fib, 1), 3), **kw)
        self._xs_all = torch.zeros((max(self.n_shell, 1), 3), **kw)
        self._v = torch.zeros((max(self.n_fib_own + self.n_shell_own + self.n_body_own, 1), 3), **kw)
        self._res = torch.zeros(max(4 * self.n_fib_own, 1), **kw)
        self._pads = {}
        for key, counts in (("fib", self.fib_counts), ("shell", self.shell_counts)):
            m = max(max(counts), 1)
            idx = torch.tensor([r * m + i for r in range(world) for i in range(counts[r])], dtype=torch.long,
                               device=device)
            self._pads[key] = (m, torch.zeros((m, 3), **kw), torch.zeros((world * m, 3), **kw), idx)

    def _gather(self, key, local, out):
        """all-gather of unequal row blocks into `out` (preallocated): pad -> all_gather_into_tensor -> index_select."""
        import torch.distributed as dist
        m, pad, gathered, idx = self._pads[key]
        if self.world == 1 or not dist.is_initialized():
            out[:local.shape[0]].copy_(local)
            return out
        pad[:local.shape[0]].copy_(local)
        dist.all_gather_into_tensor(gathered, pad, group=self.group)
        if idx.numel():
            self.torch.index_select(gathered, 0, idx, out=out[:idx.numel()])
        return out

    def _stream(self):
        t = self.torch
        return t.cuda.current_stream().cuda_stream if self._v.is_cuda else 0

    def apply(self, x_fib_own, x_shell_own, body_density, body_forces, body_torques, link_own, eta):
        """x_fib_own (4 n_fib_own,), x_shell_own (n_shell_own, 3), body_* as held by rank 0 ((n_body, 3), (n_bodies, 3)
        twice; other ranks pass buffers of the same shape to receive the broadcast), link_own (n_fibers_own, 7) or
        None.  Returns views (res_fib_own, v_shell_own, v_body_own) valid until the next call."""
        import torch.distributed as dist
        fl, st = self.flow, self._stream()
        if self.n_fib_own:
            fl.apply_fiber_force_device(x_fib_own.data_ptr(), self._fw_own.data_ptr(), st)
        self._gather("fib", self._fw_own[:self.n_fib_own], self._fw_all)
        self._gather("shell", x_shell_own, self._xs_all)
        if self.world > 1 and dist.is_initialized():
            for t in (body_density, body_forces, body_torques):
                if t.numel():
                    dist.broadcast(t, src=0, group=self.group)
        fl.matvec_device(self._fw_all.data_ptr(), self._xs_all.data_ptr(), body_density.data_ptr(),
                         body_forces.data_ptr(), body_torques.data_ptr(), float(eta), self._v.data_ptr(), st)
        if self.n_fib_own:
            fl.fiber_matvec_device(x_fib_own.data_ptr(), self._v.data_ptr(),
                                   link_own.data_ptr() if link_own is not None else 0, self._res.data_ptr(), st)
        a, b = self.n_fib_own, self.n_fib_own + self.n_shell_own
        return self._res[:4 * a], self._v[a:b], self._v[b:b + self.n_body_own]
