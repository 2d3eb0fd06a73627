#!/usr/bin/env bash
# 3xTF32 tensor-core path for fp32 corpora: parity + timing against the fp32 scan
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp7.log) 2>&1
timeout 120 python - <<'PY'
import numpy as np
import myscaledb_b200 as b2
rng=np.random.default_rng(0)
y=rng.standard_normal((20000,768)).astype(np.float32); x=rng.standard_normal((128,768)).astype(np.float32)
c=b2.Corpus(b2.IP,768).append(y); c.set_path(2)
d,i=c.search(x,10)
ref=(x.astype(np.float64)@y.T.astype(np.float64))
ri=np.argsort(-ref,axis=1)[:,:10]
print('first call ids equal frac',(ri==i).mean(),'max abs err',np.abs(np.take_along_axis(ref,i,1)-d).max())
PY
echo "rc=$?"
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -x --timeout 120 2>&1 | tail -5
timeout 300 python - <<'PY'
import numpy as np, torch, time
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
for nq in (16,64,256,1024):
    for path in (2,1):
        if path==1 and nq>256: continue
        c.set_path(path); c.search(q[:nq],10); c.kernel_time(reset=True)
        reps=5 if path==2 else 2
        for _ in range(reps): dd,ii=c.search(q[:nq],10)
        ms,kn=c.kernel_time(reset=True)
        per=ms/reps
        print(f'nq={nq} path={path} kernel ms/search {per:.2f}  launches/search {kn/reps:.0f}  eff TFLOP/s {2*nq*n*d/per/1e9:.1f}  corpus GB/s {n*d*4/(per)/1e6*(kn/reps):.0f}')
        if path==2: keep=(dd,ii)
        elif nq<=256: print('   ids equal scan vs tf32x3:', (keep[1]==ii).mean(), 'max dis diff', np.abs(keep[0]-dd).max())
PY
