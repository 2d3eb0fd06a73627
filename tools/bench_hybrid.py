#!/usr/bin/env python
"""BASELINE config 5 at its own scale, sharded: HybridSearch(fusion_type = 'RRF') over a text column and a vector column,
documents (and their vectors) split over the GPUs of one node.

    python tools/bench_hybrid.py --docs 2000000                                 # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/bench_hybrid.py --docs 10000000                                   # cfg 5: 10 M docs over 8 GPUs

One step = one batch of `--nq` hybrid queries (3 query terms + one 768-d vector each), LIMIT 10, num_candidates = 3 x LIMIT:
  text   1. every rank counts its shard's N / tokens / df(term) and ONE all-reduce(sum) makes them table-wide
            (ReadWithHybridSearch::getStatisticForTextSearch, ReadWithHybridSearch.cpp:89-209) -- b200_comm_allreduce_sum_u64;
         2. b200_bm25_search_batch scores the shard with those statistics (bm25.cu), top-30 per query;
         3. the per-shard lists meet in b200_comm_gather_merge (all-gather + merge kernel, descending);
  vector 4. b200_sharded_index_search: MSTG shard scan + all-gather + merge (top-30);
  fusion 5. rank 0: b200_hybrid_fusion_batch (RRF, k = 60) -> top-10 (HybridSearchFusion.cpp semantics, fusion.cu).
Documents are generated per 100 000-id block from the block number alone, so the table is the same however many GPUs
share it: `text_checksum` (over the merged BM25 top-30 ids of every query) must be identical for every world size -- the
script prints it so that N = 1 and N = 8 runs can be compared; scores are fp32 sums in a fixed clause order, so they are too.
The timed region includes the host tokenisation of the queries and every host<->device copy."""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myscaledb_b200 as b2  # noqa: E402
from myscaledb_b200 import search as S  # noqa: E402
from myscaledb_b200 import sharding as SH  # noqa: E402

BLOCK = 100_000
VOCAB = 100_000


def zipf():
    p = 1.0 / np.arange(1, VOCAB + 1) ** 1.1
    return p / p.sum()


def block_docs(block, pz, words):
    rng = np.random.default_rng(1000 + block)
    lens = 1 + rng.poisson(63, BLOCK)
    ids = rng.choice(VOCAB, size=int(lens.sum()), p=pz)
    pos = 0
    for i in range(BLOCK):
        yield block * BLOCK + i, " ".join(words[ids[pos:pos + lens[i]]])
        pos += lens[i]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=2_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=512)
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nprobe", type=int, default=2)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        comm = SH.Comm.from_torch_distributed(dev)
    assert a.docs % (BLOCK * world) == 0, "docs must be a multiple of 100000 x world"
    per = a.docs // world
    lo = rank * per
    k_cand = 3 * a.limit

    # ---- text shard
    pz, words = zipf(), np.array([f"t{i}" for i in range(VOCAB)])
    t0 = time.perf_counter()
    tix = b2.BM25Index(1)
    for blk in range(lo // BLOCK, (lo + per) // BLOCK):
        for doc, text in block_docs(blk, pz, words):
            tix.add_doc(doc, [text])
    tix.commit()
    t_text = time.perf_counter() - t0

    # ---- vector shard (clustered rows; ids are local, id_offset makes them table-wide)
    t0 = time.perf_counter()
    g = torch.Generator(device=dev).manual_seed(5)
    n_cent = 4096
    cents = torch.randn((n_cent, a.dim), generator=g, device=dev)
    vix = b2.VectorIndex("MSTG", b2.COSINE, a.dim, f"nprobe={a.nprobe},keep_raw=0")
    vix.reserve(per)
    gs = torch.Generator(device=dev).manual_seed(100 + rank)
    chunk = 500_000
    sample = []
    for c0 in range(0, per, chunk):
        n = min(chunk, per - c0)
        rows = cents[torch.randint(0, n_cent, (n,), generator=gs, device=dev)] + 0.3 * torch.randn((n, a.dim), generator=gs, device=dev)
        sample.append(rows[::16].clone())
        del rows
    tr = torch.cat(sample)
    torch.cuda.synchronize()
    vix.train_device(tr.data_ptr(), tr.shape[0])
    del tr, sample
    gs = torch.Generator(device=dev).manual_seed(100 + rank)
    for c0 in range(0, per, chunk):
        n = min(chunk, per - c0)
        rows = cents[torch.randint(0, n_cent, (n,), generator=gs, device=dev)] + 0.3 * torch.randn((n, a.dim), generator=gs, device=dev)
        torch.cuda.synchronize()
        vix.add_device(rows.data_ptr(), n)
        del rows
    vix.finalize()
    torch.cuda.synchronize()
    t_vec = time.perf_counter() - t0

    # ---- queries (the same on every rank)
    rng = np.random.default_rng(77)
    tail = pz[100:] / pz[100:].sum()
    sentences = [" ".join(words[100 + rng.choice(VOCAB - 100, size=3, p=tail)]) for _ in range(a.nq)]
    gq = torch.Generator(device="cpu").manual_seed(9)
    qv = (cents.cpu()[torch.randint(0, n_cent, (a.nq,), generator=gq)] + 0.3 * torch.randn((a.nq, a.dim), generator=gq)).pin_memory()
    d_q = torch.empty((a.nq, a.dim), device=dev)
    v_dis = torch.empty((a.nq, k_cand), device=dev)
    v_ids = torch.empty((a.nq, k_cand), dtype=torch.int64, device=dev)
    tstream = torch.cuda.Stream(dev)            # the sharded entry points want a real (non-NULL) stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    phase = {"stats": 0.0, "text": 0.0, "text_merge": 0.0, "vector": 0.0, "fusion": 0.0}
    uniq_terms = sorted({t for s in sentences for t in tix.query_terms(s)})

    def step(record):
        t = [time.perf_counter()]
        # 1. table-wide statistics
        local = SH.bm25_local_stats(tix, 1, uniq_terms, (0,))
        summed = comm.allreduce_sum_u64(local) if comm else local
        stats = SH.bm25_global_stats(summed, 1, uniq_terms, (0,))
        t.append(time.perf_counter())
        # 2. shard scores
        res = tix.search_batch(sentences, k_cand, stats=stats)
        t.append(time.perf_counter())
        # 3. lists of all shards -> one list per query
        td = np.full((a.nq, k_cand), -np.inf, np.float32)
        ti = np.full((a.nq, k_cand), -1, np.int64)
        for q, (rows, sc) in enumerate(res):
            td[q, :len(sc)] = sc
            ti[q, :len(sc)] = rows
        if comm:
            td, ti = comm.gather_merge_host(td, ti, descending=True)
        t.append(time.perf_counter())
        # 4. vectors
        d_q.copy_(qv, non_blocking=True)
        if comm:
            comm.sharded_index_search(vix, b2.COSINE, d_q.data_ptr(), a.nq, k_cand, "", v_dis.data_ptr(), v_ids.data_ptr(), lo, stream)
        else:
            vix.search_device(d_q.data_ptr(), a.nq, k_cand, v_dis.data_ptr(), v_ids.data_ptr(), id_offset=lo, stream=stream)
        hv_d, hv_i = v_dis.cpu().numpy(), v_ids.cpu().numpy()
        t.append(time.perf_counter())
        # 5. fusion on the initiator
        fused = None
        if rank == 0:
            f_ids, f_sc, _ = S.hybrid_fusion_arrays("rrf", hv_i, hv_d, ti, td, a.limit, fusion_k=60, vector_scan_direction=1)
            fused = (f_ids, f_sc)
        t.append(time.perf_counter())
        if record:
            for name, d in zip(phase, np.diff(t)):
                phase[name] += d
        return fused, ti

    for _ in range(a.warmup):
        step(False)
    if comm:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fused, ids = step(True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if comm:
        import torch.distributed as dist
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = {"workload": f"HybridSearch RRF, {a.docs} docs x (text ~64 tokens Zipf(1.1) vocab {VOCAB} + {a.dim}-d vector), "
                           f"batch {a.nq}, LIMIT {a.limit}, num_candidates {k_cand}, MSTG nprobe {a.nprobe}",
               "n_gpus": world, "docs_per_gpu": per, "qps": a.nq * a.steps / float(dt.item()), "ms_per_batch": float(dt.item()) / a.steps * 1e3,
               "phase_ms_rank0": {k_: v / a.steps * 1e3 for k_, v in phase.items()},
               "text_build_s_per_rank": t_text, "vector_build_s_per_rank": t_vec,
               "text_checksum": zlib.crc32(np.ascontiguousarray(ids).tobytes()),
               "text_hits": int((ids >= 0).sum()), "fused_first": [fused[0][0][:3].tolist(), fused[1][0][:3].tolist()] if fused else None,
               "timing": "host wall clock around the whole batch (tokenise + all-reduce + score + gather + vector scan + fusion), max over ranks"}
        print(json.dumps(out))
    tix.close()
    vix.close()
    if comm:
        comm.close()
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
