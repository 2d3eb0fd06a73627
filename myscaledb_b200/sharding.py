"""Host-side partitioning of a table's parts across the GPUs of one box.

The reference treats every MergeTree part as an independent search unit and merges per-part
top-k lists by score (src/VectorIndex/Storages/MergeTreeSelectWithHybridSearchProcessor.cpp:1149-1241
-> MergeTreeBaseSearchManager.cpp:207-299).  Here a part (or an equal row range of a big part)
is pinned to one GPU; queries are broadcast; the only exchange is one all-gather of
[nq][k] (distance fp32, global row id int64) followed by one merge kernel.
"""
from __future__ import annotations


def shard_range(total_rows: int, world: int, rank: int, align: int = 1):
    """Row range [start, stop) of `rank`: equal shares rounded to `align`, remainder to the last ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    units = total_rows // align
    base, extra = divmod(units, world)
    start_u = rank * base + min(rank, extra)
    stop_u = start_u + base + (1 if rank < extra else 0)
    start, stop = start_u * align, stop_u * align
    if rank == world - 1:
        stop = total_rows
    return start, stop


def gather_layout(world: int, nq: int, k: int):
    """Shapes of the all-gather buffers consumed by b200_topk_merge_device."""
    return (world, nq, k)


def assign_parts(part_rows, world: int):
    """Greedy size-balanced assignment of parts to GPUs (largest first). Returns list of lists of part indices."""
    order = sorted(range(len(part_rows)), key=lambda i: -part_rows[i])
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        g = min(range(world), key=lambda j: load[j])
        out[g].append(i)
        load[g] += part_rows[i]
    return out
