"""CPU tests of the host-side planners (no GPU): the launch plan of the plain pair kernel and the work list of the
symmetric kernel, through the C ABI (skb_plan_query / skb_sym_plan_query)."""
import pytest

from skellysim_b200 import capi

SRC_TILE = 128


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n_trg,n_src", [(1, 1), (6, 96000), (512, 512), (743, 1229), (8000, 32000), (40000, 32000),
                                         (102400, 96000), (1000000, 1000000), (2829, 90496)])
def test_launch_plan_covers_the_problem(kind, n_trg, n_src):
    p = capi.plan_query(kind, n_trg, n_src)
    assert p["T"] in (1, 2, 4, 8)
    tile_t = 128 * p["T"]
    assert p["grid_x"] == -(-n_trg // tile_t)                       # every target in exactly one tile
    n_src_tiles = -(-n_src // SRC_TILE)
    assert p["n_splits"] >= 1 and p["tiles_per_split"] >= 1
    assert p["n_splits"] * p["tiles_per_split"] >= n_src_tiles      # every source tile in some split
    assert (p["n_splits"] - 1) * p["tiles_per_split"] < n_src_tiles  # no empty split
    assert p["n_splits"] <= 256


def test_launch_plan_fills_the_gpu_for_few_targets():
    # few targets, many sources (listener path / remainder targets): the source dimension must be split
    p = capi.plan_query(0, 6, 96000)
    assert p["grid_x"] == 1 and p["n_splits"] >= 100
    # plenty of targets: no need for many splits
    p = capi.plan_query(0, 1000000, 1000000)
    assert p["grid_x"] * p["n_splits"] >= 148 * 2


def test_forced_tuning_is_honoured():
    p = capi.plan_query(0, 50000, 50000, force_T=2, force_S=5)
    assert p["T"] == 2 and p["n_splits"] == 5


@pytest.mark.parametrize("nb", [2, 3, 32, 63, 94, 188, 977])
@pytest.mark.parametrize("parts", [1, 2, 8])
def test_symmetric_work_list_covers_the_upper_triangle_exactly_once(nb, parts):
    gpb = capi.sym_groups_per_block()
    covered = {}   # block row -> sorted list of group ranges
    for part in range(parts):
        items, row_begin = capi.sym_plan_query(nb, part, parts)
        slots = sorted(it[3] for it in items)
        assert slots == list(range(len(items)))                    # forward-partial slabs: a permutation
        assert row_begin[0] == 0 and row_begin[-1] == len(items)
        sizes = [g1 - g0 for (_, g0, g1, _) in items]
        assert sizes == sorted(sizes, reverse=True)                # launch order: large items first
        by_slot = {it[3]: it for it in items}
        for b in range(nb):
            for s in range(row_begin[b], row_begin[b + 1]):
                assert by_slot[s][0] == b                          # row b's slabs are contiguous
        for (i, g0, g1, _) in items:
            assert 0 <= i and gpb * i <= g0 < g1 <= gpb * nb        # block i's own groups (the diagonal) and beyond
            assert g0 % 4 == 0                                     # items start on a stage boundary
            covered.setdefault(i, []).append((g0, g1))
    total = 0
    for i in range(nb):
        rs = sorted(covered.get(i, []))
        assert rs[0][0] == gpb * i and rs[-1][1] == gpb * nb        # the whole row, diagonal block included ...
        for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
            assert a1 == b0                                        # ... exactly once
        total += rs[-1][1] - rs[0][0]
    assert total == gpb * nb * (nb + 1) // 2


@pytest.mark.parametrize("nb,parts", [(32, 1), (94, 1), (94, 8), (266, 8), (977, 8)])
def test_symmetric_work_list_has_a_fine_tail(nb, parts):
    # large items first, the last few per cent of the work in single stages: every CTA slot drains within ~one stage
    items, _ = capi.sym_plan_query(nb, 0, parts)
    sizes = [g1 - g0 for (_, g0, g1, _) in items]
    work = sum(sizes)
    slots = 148 * 2   # resident CTAs of the kernel (2 per SM with 8 nodes per thread)
    assert sizes[-1] <= 4
    tail = [s for s in sizes if s <= 4]
    assert sum(tail) >= min(0.04 * work, 4 * 4 * slots) - 16             # enough fine items to level the slots ...
    assert len(tail) >= min(slots, len(sizes)) or work < 8 * slots
    assert max(sizes) <= max(4, 0.4 * work / slots + 4)            # ... and no item is a big share of a slot


def test_symmetric_work_list_is_balanced_across_parts():
    loads = []
    for part in range(8):
        items, _ = capi.sym_plan_query(178, part, 8)
        loads.append(sum(g1 - g0 for (_, g0, g1, _) in items))
    assert max(loads) - min(loads) <= 0.03 * max(loads) + 8 * capi.sym_groups_per_block()


# ---- skb_partition_query: which rows of [fibers | periphery | bodies] a group member owns (host-only) ----------------
import numpy as np  # noqa: E402
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(nodes=st.lists(st.sampled_from([8, 16, 24, 32, 48, 64, 96, 128]), min_size=0, max_size=200),
       n_shell=st.integers(0, 5000), n_body=st.integers(0, 900), members=st.integers(1, 16))
def test_row_partition_tiles_every_class_with_whole_fibers(nodes, n_shell, n_body, members):
    """Every member owns a contiguous run of WHOLE fibers, of periphery rows and of body rows; the runs tile each class
    in member order with no gap and no overlap; fibers are balanced by node count (what the pair work scales with)."""
    from skellysim_b200 import capi
    parts = [capi.partition_query(nodes, n_shell, n_body, members, m) for m in range(members)]
    for cls, total in ((0, len(nodes)), (2, n_shell), (4, n_body)):
        assert parts[0][cls] == 0 and parts[-1][cls + 1] == total
        for a, b in zip(parts, parts[1:]):
            assert a[cls] <= a[cls + 1] == b[cls]
    if nodes:
        off = np.concatenate([[0], np.cumsum(nodes)])
        own = [off[p[1]] - off[p[0]] for p in parts]
        # no member is further from the ideal share than the largest fiber
        assert max(abs(o - off[-1] / members) for o in own) <= max(nodes) + 1e-9
