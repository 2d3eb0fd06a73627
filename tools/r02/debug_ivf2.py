import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import myscaledb_b200 as b2
def recall(a, b):
    return float(np.mean([len(set(x.tolist()) & set(y.tolist())) / len(y) for x, y in zip(a, b)]))
rng = np.random.default_rng(1)
d, n, nc, nl = 96, 200000, 500, 256
centres = rng.standard_normal((nc, d)).astype(np.float32)
y = (centres[rng.integers(0, nc, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
q = (centres[rng.integers(0, nc, 64)] + 0.3 * rng.standard_normal((64, d))).astype(np.float32)
flat = b2.Corpus(b2.L2, d).append(y); dt, it = flat.search(q, 10); flat.close()
ix = b2.VectorIndex("MSTG", b2.L2, d, f"ncentroids={nl}").build(y)
for par in ("nprobe=1", "nprobe=8", "nprobe=8, pages_per_chunk=1", "nprobe=8, pages_per_chunk=2", "nprobe=8, pages_per_chunk=16", "nprobe=8, refine_factor=1", "nprobe=1, pages_per_chunk=16"):
    r = []
    for rep in range(3):
        dg, ig = ix.search(q, 10, par); r.append(round(recall(ig, it), 4))
    print(os.environ.get("B200_IVF_COOP", "default"), par, "recall x3", r, flush=True)
