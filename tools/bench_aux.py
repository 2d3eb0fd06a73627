#!/usr/bin/env python3
"""Secondary measurements (not the driver's bench line): IVFPQ, the two-stage (MSTG-type) index and
BM25 at moderate single-GPU scale, shaped after BASELINE.json configs 3-5.  Prints one JSON line per
workload; results are pasted into DESIGN.md section 7.  Usage: python tools/bench_aux.py [ivfpq] [mstg] [bm25]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myscaledb_b200 as b2  # noqa: E402
import oracle as orc  # noqa: E402
from myscaledb_b200 import search as S  # noqa: E402


def clustered(n, d, n_centres, seed, spread=0.3, nq=1024):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_centres, d)).astype(np.float32)
    y = np.empty((n, d), np.float32)
    step = 500_000
    for i in range(0, n, step):
        m = min(step, n - i)
        y[i:i + m] = centres[rng.integers(0, n_centres, m)] + spread * rng.standard_normal((m, d)).astype(np.float32)
    q = centres[rng.integers(0, n_centres, nq)] + spread * rng.standard_normal((nq, d)).astype(np.float32)
    return y, q.astype(np.float32)


def recall(ids, truth):
    return float(np.mean([len(set(a.tolist()) & set(b.tolist())) / len(b) for a, b in zip(ids, truth)]))


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def bench_ivfpq():
    n, d, nlist, m, nq, k = 5_000_000, 96, 4096, 96, 10_000, 10
    y, q = clustered(n, d, 10_000, 7, nq=nq)
    y /= np.linalg.norm(y, axis=1, keepdims=True)  # Deep1B shape: unit vectors
    t0 = time.perf_counter()
    ix = b2.VectorIndex("IVFPQ", b2.L2, d, f"ncentroids={nlist}, M={m}").build(y)
    t_build = time.perf_counter() - t0
    flat = b2.Corpus(b2.L2, d).append(y)
    _, truth = flat.search(q[:1000], k)
    out = {"workload": f"IVFPQ nlist={nlist} m={m}, {n} x {d}-d fp32 unit vectors, batch {nq}, top-{k} (config 4 shape, 1 GPU)",
           "build_s": t_build, "points": []}
    for nprobe in (8, 32, 64):
        t, (dis, ids) = timed(lambda: ix.search(q, k, f"nprobe={nprobe}, exact_batch=0"))
        out["points"].append({"nprobe": nprobe, "qps": nq / t, "recall@10_first_stage": recall(ids[:1000], truth),
                              "code_bytes_per_query": nprobe / nlist * n * m,
                              "code_GB_per_s": nprobe / nlist * n * m * nq / t / 1e9})
    # the index's own batch planner (no override): one exact 3xTF32 pass over the raw fp32 rows when that is cheaper
    t, (dis, ids) = timed(lambda: ix.search(q, k, "nprobe=32"))
    out["planner_auto"] = {"qps": nq / t, "recall@10": recall(ids[:1000], truth), "candidates": ix.last_num_candidates}
    print(json.dumps(out))


def bench_mstg():
    n, d, nq, k = 2_000_000, 768, 256, 10
    y, q = clustered(n, d, 10_000, 5, nq=nq)
    t0 = time.perf_counter()
    ix = b2.VectorIndex("MSTG", b2.L2, d, "ncentroids=2048, M=96").build(y)
    t_build = time.perf_counter() - t0
    flat = b2.Corpus(b2.L2, d).append(y)
    t_flat, (_, truth) = timed(lambda: flat.search(q, k), reps=2)
    out = {"workload": f"two-stage (MSTG-type: IVFPQ + exact refine) {n} x {d}-d fp32, batch {nq}, top-{k} (config 3 shape, 1 GPU shard)",
           "build_s": t_build, "exact_flat_scan_qps": nq / t_flat, "points": []}
    for nprobe, rf in ((16, 8), (32, 16), (64, 16)):
        t, (dis, ids) = timed(lambda: ix.search(q, k, f"nprobe={nprobe}, refine_factor={rf}, exact_batch=0"))
        out["points"].append({"nprobe": nprobe, "refine_factor": rf, "qps": nq / t, "recall@10": recall(ids, truth)})
    for nq_b in (256, 16, 1):
        t, (dis, ids) = timed(lambda: ix.search(q[:nq_b], k, "nprobe=32, refine_factor=16"))
        out.setdefault("planner_auto", []).append({"batch": nq_b, "qps": nq_b / t, "recall@10": recall(ids, truth[:nq_b]),
                                                   "candidates": ix.last_num_candidates})
    # CPU arm: the oracle's exact threaded brute force on a row sample (the reference would run its closed CPU MSTG)
    rows = 100_000
    t0 = time.perf_counter()
    orc.knn_flat_parts_blas(orc.L2, q, y[:rows], k, min(64, os.cpu_count() or 8)) or orc.knn_flat_parts(orc.L2, q, y[:rows], k, os.cpu_count() or 8)
    out["cpu_exact_qps_scaled"] = nq / ((time.perf_counter() - t0) * n / rows)
    print(json.dumps(out))


def bench_bm25():
    n_docs, vocab, nq = 1_000_000, 100_000, 512
    rng = np.random.default_rng(9)
    pz = 1.0 / np.arange(1, vocab + 1) ** 1.1
    pz /= pz.sum()
    words = np.array([f"t{i}" for i in range(vocab)])
    lens = 1 + rng.poisson(63, n_docs)
    flat_ids = rng.choice(vocab, size=int(lens.sum()), p=pz)
    g = b2.BM25Index(1)
    o = orc.BM25Index(1)
    t0 = time.perf_counter()
    pos = 0
    for dno in range(n_docs):
        text = " ".join(words[flat_ids[pos:pos + lens[dno]]])
        pos += lens[dno]
        g.add_doc(dno, [text])
        if dno < 200_000:
            o.add_doc(dno, [text])
    g.commit()
    t_build = time.perf_counter() - t0
    queries = [" ".join(words[100 + rng.choice(vocab - 100, size=3, p=pz[100:] / pz[100:].sum())]) for _ in range(nq)]
    t, res = timed(lambda: g.search_batch(queries, 30))
    tm = g.last_timing()
    post = sum(g.doc_freq(t_) for qs in queries for t_ in set(qs.split()))
    t0 = time.perf_counter()
    for qs in queries[:64]:
        o.search(qs, 30)
    t_cpu = (time.perf_counter() - t0) / 64 * (n_docs / 200_000)
    print(json.dumps({"workload": f"BM25 top-30 (num_candidates = 3 x LIMIT 10), {n_docs} docs, Zipf(1.1) vocab {vocab}, "
                                  f"len ~ Poisson(64), batch {nq} x 3 terms (config 5 text side, 1 GPU)",
                      "build_s_host_tokenise_and_upload": t_build, "qps": nq / t, "postings_scored_per_batch": post,
                      "posting_GB_per_s": post * 9 / t / 1e9, "python_call_ms": t * 1e3, "c_call_ms": tm["call_ms"],
                      "score_kernel_ms": tm["kernel_ms"], "score_kernel_posting_GB_per_s": tm["postings"] * 9 / (tm["kernel_ms"] * 1e-3) / 1e9 if tm["kernel_ms"] else None,
                      "cpu_oracle_qps_1_thread_scaled": 1.0 / t_cpu}))


def bench_flat10k():
    """BASELINE.json configs[0]: FLAT L2 distance(), 10k rows x 128-d fp32, single query, top-10."""
    rng = np.random.default_rng(1)
    y = rng.standard_normal((10_000, 128)).astype(np.float32)
    q = np.random.default_rng(2).standard_normal((1, 128)).astype(np.float32)
    t_cold, _ = timed(lambda: b2.part_scan(b2.L2, q, y, 10), reps=50)          # host part column in, results out
    c = b2.Corpus(b2.L2, 128).append(y)
    c.enable_timing(True)
    t_res, _ = timed(lambda: c.search(q, 10), reps=200)                       # resident part / FLAT index
    kms, kn = c.kernel_time(reset=True)
    t_cpu, _ = timed(lambda: orc.part_scan(orc.L2, q, y, 10), reps=50)
    print(json.dumps({"workload": "FLAT L2 distance(), 10k x 128 fp32, nq=1, top-10 (config 1; 5.12 MB per query)",
                      "gpu_part_scan_host_buffers_us": t_cold * 1e6, "gpu_resident_search_us": t_res * 1e6,
                      "scan_kernel_us": kms / max(kn, 1) * 1e3, "scan_kernel_GB_per_s": 5.12e6 / (kms / max(kn, 1) * 1e-3) / 1e9,
                      "cpu_oracle_1_thread_us": t_cpu * 1e6,
                      "note": "launch/latency-bound: one 5 MB part is smaller than one wave of loads"}))


def bench_ingest():
    """Cold brute force: the part column starts in pageable host memory (the reference's vectorScanWithoutIndex case)."""
    n, d = 1_000_000, 768
    rng = np.random.default_rng(3)
    y = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((1, d), dtype=np.float32)
    out = {"workload": f"cold FLAT L2 part scan, {n} x {d} fp32 part in pageable host memory ({n * d * 4 / 1e9:.2f} GB), nq=1, top-10",
           "points": []}
    for threads in ("0", "2", "4", "6", "8"):
        os.environ["B200_INGEST_THREADS"] = threads
        t, _ = timed(lambda: b2.part_scan(b2.L2, q, y, 10), reps=3)
        out["points"].append({"ingest_threads": int(threads), "s_per_scan": t, "host_to_hbm_GB_per_s": n * d * 4 / t / 1e9})
    del os.environ["B200_INGEST_THREADS"]
    t0 = time.perf_counter()
    orc.part_scan(orc.L2, q, y[:200_000], 10)
    out["cpu_oracle_1_thread_s_scaled"] = (time.perf_counter() - t0) * n / 200_000
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ivfpq", "mstg", "bm25"]
    for w in which:
        {"ivfpq": bench_ivfpq, "mstg": bench_mstg, "bm25": bench_bm25, "flat10k": bench_flat10k, "ingest": bench_ingest}[w]()
