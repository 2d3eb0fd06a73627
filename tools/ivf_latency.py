#!/usr/bin/env python3
"""Single-query / small-batch latency of the two-stage index (run under `ncu -k regex:...` for the per-kernel split)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myscaledb_b200 as b2  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_aux import clustered, recall, timed  # noqa: E402

n, d = 1_000_000, 768
y, q = clustered(n, d, 5000, 5, nq=64)
ix = b2.VectorIndex("MSTG", b2.L2, d, "ncentroids=1024, M=96").build(y)
flat = b2.Corpus(b2.L2, d).append(y)
_, truth = flat.search(q, 10)
for nq in (1, 4, 16):
    for params in ("nprobe=32, refine_factor=16, exact_batch=0", "nprobe=32, refine_factor=16, exact_batch=1"):
        t, (dis, ids) = timed(lambda: ix.search(q[:nq], 10, params), reps=int(os.environ.get("REPS", "20")))
        print(f"nq={nq} {params}: {t * 1e6:.0f} us per call, recall {recall(ids, truth[:nq]):.3f}", flush=True)
