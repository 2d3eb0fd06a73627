#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp17.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python - <<'PY'
import numpy as np, torch
import myscaledb_b200 as b2
from myscaledb_b200 import search as S
n,d=4_000_000,768
dev=torch.device('cuda:0')
g=torch.Generator(device=dev); g.manual_seed(1)
q=np.random.default_rng(1).standard_normal((1024,d)).astype(np.float32)
for dt,name in ((S.F32,'fp32'),(S.BF16,'bf16')):
    corpus=torch.empty((n,d),dtype=torch.float32 if dt==S.F32 else torch.bfloat16,device=dev)
    for off in range(0,n,500_000):
        corpus[off:off+500_000]=torch.randn((500_000,d),generator=g,device=dev).to(corpus.dtype)
    c=b2.Corpus(b2.IP,d,dtype=dt); c.adopt_device(corpus.data_ptr(),n)
    c.enable_timing(True)
    for nq in (16,128,129,300,1000,1024):
        c.search(q[:nq],10); c.kernel_time(reset=True)
        for _ in range(5): dd,ii=c.search(q[:nq],10)
        ms,kn=c.kernel_time(reset=True)
        print(f'{name} nq={nq}: kernel ms/search {ms/5:.2f} eff TFLOP/s {2*nq*n*d/(ms/5)/1e9:.1f}')
    c.close(); del corpus
PY
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line.json; cat gpurun_out/bench_line.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms_per_step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'],[round(x['GB_per_s']) for x in d['flat_scan']], round(d['fp32_batch']['ms_per_launch'],2), d['cpu_baseline']['value'])"
