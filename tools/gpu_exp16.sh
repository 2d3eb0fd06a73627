#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp16.log) 2>&1
timeout 600 python -m pytest tests/test_gpu_flat.py tests/test_gpu_edges.py tests/test_gpu_hybrid.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python - <<'PY'
import numpy as np, torch
import myscaledb_b200 as b2
n,d=4_000_000,768
dev=torch.device('cuda:0')
corpus=torch.empty((n,d),dtype=torch.float32,device=dev)
g=torch.Generator(device=dev); g.manual_seed(1)
for off in range(0,n,500_000):
    corpus[off:off+500_000]=torch.randn((500_000,d),generator=g,device=dev)
c=b2.Corpus(b2.IP,d); c.adopt_device(corpus.data_ptr(),n)
q=np.random.default_rng(1).standard_normal((1024,d)).astype(np.float32)
c.enable_timing(True)
for nq in (1,8,16,64,128,129,256,1024):
    c.search(q[:nq],10); c.kernel_time(reset=True)
    for _ in range(5): dd,ii=c.search(q[:nq],10)
    ms,kn=c.kernel_time(reset=True)
    print(f'nq={nq} auto path: kernel ms/search {ms/5:.2f} launches/search {kn/5:.0f} eff TFLOP/s {2*nq*n*d/(ms/5)/1e9:.1f}')
PY
