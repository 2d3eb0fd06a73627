import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import myscaledb_b200 as b2
from myscaledb_b200 import search as S

def recall(a, b):
    return float(np.mean([len(set(x.tolist()) & set(y.tolist())) / len(y) for x, y in zip(a, b)]))

rng = np.random.default_rng(1)
for d, n, nc, nl in ((96, 200000, 500, 256), (256, 200000, 500, 256), (768, 200000, 500, 256), (768, 200000, 500, 1024)):
    centres = rng.standard_normal((nc, d)).astype(np.float32)
    y = (centres[rng.integers(0, nc, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    q = (centres[rng.integers(0, nc, 64)] + 0.3 * rng.standard_normal((64, d))).astype(np.float32)
    flat = b2.Corpus(b2.L2, d).append(y)
    dt, it = flat.search(q, 10)
    flat.close()
    ix = b2.VectorIndex("MSTG", b2.L2, d, f"ncentroids={nl}").build(y)
    sz = ix.list_sizes()
    print(f"d={d} n={n} nlist={nl}: list sizes min {sz.min()} max {sz.max()} nonempty {(sz>0).sum()} sum {sz.sum()}")
    for par in ("nprobe=1", "nprobe=8", f"nprobe={nl}", f"nprobe={nl}, refine_factor=1", "nprobe=8, pages_per_chunk=1"):
        dg, ig = ix.search(q, 10, par)
        print("   ", par, "recall", round(recall(ig, it), 4), "first row", ig[0, :4], it[0, :4], dg[0, :2], dt[0, :2])
    d1, i1 = ix.search(q, 40, f"nprobe={nl}", first_stage_only=True)
    print("    first stage k=40 contains truth:", round(float(np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(i1, it)])), 4))
    # device-side streamed build, like tools/bench_ivf.py
    ty = torch.tensor(y).cuda()
    ix2 = b2.VectorIndex("MSTG", b2.L2, d, f"ncentroids={nl}").reserve(n)
    samp = ty[::3].contiguous()
    ix2.train_device(samp.data_ptr(), samp.shape[0])
    for off in range(0, n, 70000):
        part = ty[off:off + 70000].contiguous()
        torch.cuda.synchronize()
        ix2.add_device(part.data_ptr(), part.shape[0])
    ix2.finalize()
    tq = torch.tensor(q).cuda(); od = torch.empty((64, 10), device="cuda"); oi = torch.empty((64, 10), dtype=torch.int64, device="cuda")
    ix2.search_device(tq.data_ptr(), 64, 10, od.data_ptr(), oi.data_ptr(), "nprobe=8", stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("    device build + search_device nprobe=8 recall", round(recall(oi.cpu().numpy(), it), 4), "sizes", ix2.list_sizes().max())
