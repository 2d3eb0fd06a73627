#!/usr/bin/env python3
"""BASELINE configs[0]: FLAT L2 distance(), 10k x 128 fp32, ONE query, top-10 -- latency per call through the C ABI.
  resident   : b200_corpus_search() on a part kept in HBM (the single-launch fused path)
  one_shot   : b200_part_scan() with the part in host memory (H2D of the 5 MB part inside the call)
  cpu_simd   : oracle/cpu_baseline.c::orc_knn_flat_simd (faiss' nx < 20 form: exact differences, AVX-512 via target_clones,
               one thread per part like the reference) -- and the scalar checker for comparison
Also the same for a 1 M-row part (where the GPU's bandwidth starts to matter)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import myscaledb_b200 as b2
import oracle as orc


def med_us(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[int(len(ts) * 0.95)] * 1e6


def main():
    out = []
    for n, d in ((10_000, 128), (1_000_000, 128)):
        rng = np.random.default_rng(1)
        y = rng.standard_normal((n, d)).astype(np.float32)
        x = rng.standard_normal((1, d)).astype(np.float32)
        c = b2.Corpus(b2.L2, d).append(y)
        dg, ig = c.search(x, 10)
        do, io = orc.knn_flat(orc.L2, x, y, 10)
        assert (ig == io).all() and np.allclose(dg, do, rtol=1e-4), "fused path disagrees with the oracle"
        c.enable_timing(True)
        res, p95 = med_us(lambda: c.search(x, 10))
        kms, kn = c.kernel_time(reset=True)
        os.environ["B200_FUSED_SCAN"] = "0"
        c2 = b2.Corpus(b2.L2, d).append(y)
        staged, _ = med_us(lambda: c2.search(x, 10))
        os.environ.pop("B200_FUSED_SCAN")
        reps = 200 if n <= 10_000 else 20
        one, _ = med_us(lambda: b2.part_scan(b2.L2, x, y, 10), reps=reps, warm=5)
        simd, _ = med_us(lambda: orc.knn_flat_simd(orc.L2, x, y, 10), reps=reps, warm=3)
        scalar, _ = med_us(lambda: orc.knn_flat(orc.L2, x, y, 10), reps=max(5, reps // 10), warm=1)
        out.append({"workload": f"FLAT L2 distance(), {n} x {d} fp32, 1 query, top-10", "resident_call_us_median": round(res, 1),
                    "resident_call_us_p95": round(p95, 1), "kernel_us": round(kms / max(kn, 1) * 1e3, 1),
                    "resident_call_us_staged_path_r1": round(staged, 1), "one_shot_part_scan_us": round(one, 1),
                    "cpu_simd_1thread_us": round(simd, 1), "cpu_scalar_checker_us": round(scalar, 1),
                    "bytes": n * d * 4, "resident_GB_per_s": round(n * d * 4 / res / 1e3, 1)})
        c.close(); c2.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
