"""N > 1 host logic on CPU: world_size-2 gloo processes shard a corpus, search their shard
(with the oracle standing in for the GPU kernel), all-gather [nq][k] lists and merge -- the result
must equal the unsharded search.  Mirrors bench.py's multi-GPU data flow."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from myscaledb_b200.sharding import assign_parts, shard_range
    for total, world, align in [(10_000_000, 8, 250_000), (1001, 3, 1), (7, 8, 1), (10_000_000, 1, 250_000)]:
        rs = [shard_range(total, world, r, align) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    a = assign_parts([10, 1, 7, 3, 3, 8], 2)
    assert sorted(sum(a, [])) == list(range(6))
    assert abs(sum([10, 1, 7, 3, 3, 8][i] for i in a[0]) - 16) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle as orc
    from myscaledb_b200.sharding import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((5000, 32)).astype(np.float32)
    x = rng.standard_normal((9, 32)).astype(np.float32)
    k = 7
    lo, hi = shard_range(len(y), world, rank)
    d, i = orc.knn_flat(orc.IP, x, y[lo:hi], k)
    i = np.where(i >= 0, i + lo, -1)
    gd = [torch.empty((9, k)) for _ in range(world)]
    gi = [torch.empty((9, k), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gd, torch.from_numpy(d))
    dist.all_gather(gi, torch.from_numpy(i))
    if rank == 0:
        D = torch.stack(gd).numpy(); I = torch.stack(gi).numpy()
        out_i = np.empty((9, k), np.int64)
        for qi in range(9):
            sc = D[:, qi].reshape(-1); lab = I[:, qi].reshape(-1)
            order = np.lexsort((lab, -sc))[:k]
            out_i[qi] = lab[order]
        _, ref = orc.knn_flat(orc.IP, x, y, k)
        q.put(bool((out_i == ref).all()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    ok = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert ok and all(p.exitcode == 0 for p in ps)


def _hybrid_worker(rank, world, port, q):
    """Two ranks hold 10 of the reference's 20 golden documents each (test 00041's two parts): one all-reduce of the
    BM25 counters, per-shard scoring with the table-wide statistics, all-gather + merge of text and vector lists,
    fusion on rank 0 -> the reference's single-part goldens."""
    sys.path.insert(0, ROOT)
    import json

    import torch
    import torch.distributed as dist

    import oracle as orc
    from myscaledb_b200.sharding import bm25_global_stats, bm25_local_stats, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))
    docs = g["00040_hybrid"]["docs"]
    lo, hi = shard_range(len(docs), world, rank)
    ix = orc.BM25Index(1)
    for rid, _, text in docs[lo:hi]:
        ix.add_doc(rid, [text])
    terms = orc.BM25Index.query_terms("Ancient")
    counters = torch.tensor(bm25_local_stats(ix, 1, terms, (0,)), dtype=torch.int64)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    stats = bm25_global_stats(counters.tolist(), 1, terms, (0,))
    k = 5
    rows, sc = ix.search("Ancient", k, stats=stats)
    t_sc = torch.full((k,), -1.0); t_id = torch.full((k,), -1, dtype=torch.int64)
    t_sc[:len(sc)] = torch.from_numpy(np.asarray(sc, np.float32)); t_id[:len(rows)] = torch.from_numpy(np.asarray(rows, np.int64))
    y = np.repeat(np.arange(lo, hi, dtype=np.float32)[:, None], 3, axis=1)
    d, i = orc.part_scan(orc.L2, np.array([[1, 1, 1]], np.float32), y, k)
    v_d = torch.from_numpy(d[0].copy()); v_i = torch.from_numpy(np.where(i[0] >= 0, i[0] + lo, -1))
    bufs = {n: [torch.empty_like(t) for _ in range(world)] for n, t in (("ts", t_sc), ("ti", t_id), ("vd", v_d), ("vi", v_i))}
    for n, t in (("ts", t_sc), ("ti", t_id), ("vd", v_d), ("vi", v_i)):
        dist.all_gather(bufs[n], t)
    if rank == 0:
        F32 = np.float32
        ts = torch.cat(bufs["ts"]).numpy(); ti = torch.cat(bufs["ti"]).numpy()
        keep = ti >= 0
        s_, p_, l_ = orc.merge_parts(ts[keep], np.repeat(np.arange(world), k)[keep], ti[keep], k, desc=True)
        text = [[int(l_[j]), float(s_[j])] for j in range(len(s_))]
        ok_text = text == [[e[0], float(F32(e[1]))] for e in g["00041_multi_parts"]["text_1part"]]
        vd = torch.cat(bufs["vd"]).numpy(); vi = torch.cat(bufs["vi"]).numpy()
        keep = vi >= 0
        s2, p2, l2 = orc.merge_parts(vd[keep], np.repeat(np.arange(world), k)[keep], vi[keep], k, desc=False)
        vec = [(0, 0, int(l2[j]), float(s2[j])) for j in range(len(s2))]
        txt = [(0, 0, a, b) for a, b in text]
        rsf = orc.hybrid_fusion("rsf", vec, txt, 10)
        got = sorted(((r[2], float(F32(r[3]))) for r in rsf), key=lambda t: (-t[1], t[0]))[:5]
        ok_rsf = [[a, b] for a, b in got] == [[e[0], float(F32(e[1]))] for e in g["00041_multi_parts"]["rsf_1part"]]
        q.put((ok_text, ok_rsf, stats["total_docs"], stats["doc_freq"][(0, terms[0])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_hybrid_search_with_global_bm25_statistics():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_hybrid_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    ok_text, ok_rsf, n_docs, df = q.get(timeout=180)
    [p.join(60) for p in ps]
    assert (n_docs, df) == (20, 2)           # table-wide statistics after the all-reduce
    assert ok_text and ok_rsf and all(p.exitcode == 0 for p in ps)
