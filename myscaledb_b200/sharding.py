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


# ----------------------------------------------------------------------------------------------
# Text side: BM25 scores must not depend on how documents are spread over GPUs.  The reference sums
# total_num_docs / total_num_tokens / doc_freq over parts before scoring
# (ReadWithHybridSearch::getStatisticForTextSearch, src/VectorIndex/Processors/ReadWithHybridSearch.cpp:89-209,
# BM25InfoInDataParts.cpp:40-94); across GPUs that is ONE all-reduce(sum) of a small int64 vector per batch.
# ----------------------------------------------------------------------------------------------
def bm25_stats_layout(n_fields: int, query_terms, fields):
    """Order of the counters in the exchanged vector: [total_docs, tokens(field 0..n_fields-1), df(field, term) ...]."""
    keys = [("docs",)] + [("tokens", f) for f in range(n_fields)]
    keys += [("df", f, t) for f in fields for t in query_terms]
    return keys


def bm25_local_stats(index, n_fields: int, query_terms, fields):
    """int64 counters of ONE shard's BM25 index (anything with total_docs / total_tokens(f) / doc_freq(t, f))."""
    total_docs = index.total_docs() if callable(index.total_docs) else index.total_docs
    out = [int(total_docs)] + [int(index.total_tokens(f)) for f in range(n_fields)]
    out += [int(index.doc_freq(t, f)) for f in fields for t in query_terms]
    return out


def bm25_global_stats(summed, n_fields: int, query_terms, fields):
    """Summed counter vector -> the `stats` dict BM25Index.search(..., stats=) takes (table-wide N, tokens, df)."""
    summed = [int(v) for v in summed]
    stats = {"total_docs": summed[0], "total_tokens": {f: summed[1 + f] for f in range(n_fields)}, "doc_freq": {}}
    pos = 1 + n_fields
    for f in fields:
        for t in query_terms:
            stats["doc_freq"][(f, t)] = summed[pos]
            pos += 1
    return stats
