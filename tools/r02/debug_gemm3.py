import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import myscaledb_b200 as b2
from myscaledb_b200 import search as S
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
n, d, nq, k = 2_000_000, 768, 256, 10
centres = torch.randn((2000, d), generator=g, device=dev)
y = (centres[torch.randint(0, 2000, (n,), generator=g, device=dev)] + 0.3 * torch.randn((n, d), generator=g, device=dev)).contiguous()
q = (centres[torch.randint(0, 2000, (nq,), generator=g, device=dev)] + 0.3 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
s = torch.cuda.current_stream().cuda_stream
def run(c, nqq, path, off=0):
    od = torch.empty((nqq, k), device=dev); oi = torch.empty((nqq, k), dtype=torch.int64, device=dev)
    c.set_path(path)
    c.search_device(q.data_ptr(), nqq, k, od.data_ptr(), oi.data_ptr(), id_offset=off, stream=s)
    torch.cuda.synchronize()
    return od.cpu().numpy(), oi.cpu().numpy(), c.last_variant()
for metric, name in ((b2.L2, "L2"), (b2.IP, "IP")):
    full = b2.Corpus(metric, d); full.adopt_device(y.data_ptr(), n)
    dA, iA, vA = run(full, nq, 0)
    dS, iS, vS = run(full, 8, 1)
    d1, i1, v1 = run(full, 128, 0)
    parts_d, parts_i = [], []
    for off in range(0, n, 500_000):
        c = b2.Corpus(metric, d); c.adopt_device(y[off:off + 500_000].data_ptr(), 500_000)
        dd, ii, vv = run(c, nq, 0, off); parts_d.append(dd); parts_i.append(ii); c.close()
    pd_, pi_ = np.concatenate(parts_d, 1), np.concatenate(parts_i, 1)
    o = np.argsort(pd_ if metric == b2.L2 else -pd_, axis=1)[:, :k]
    iB = np.take_along_axis(pi_, o, 1)
    same = lambda a, b: float(np.mean([len(set(x.tolist()) & set(z.tolist())) / k for x, z in zip(a, b)]))
    print(name, "full(256q)", vA, "vs scan(8q)", same(iA[:8], iS), "| chunks vs scan", same(iB[:8], iS), "| full(128q)", v1, "vs scan", same(i1[:8], iS),
          "| full vs chunks", same(iA, iB))
    print("   scan ids", iS[0][:5], dS[0][:3], "full ids", iA[0][:5], dA[0][:3], "chunk ids", iB[0][:5])
    full.close()
