#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu16.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv
echo "== 2-rank NCCL test"
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 1000 2>&1 | tail -15 | cut -c1-400
echo "== bench N=2 (10M rows headline + 20M-row index)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 --index-rows 20000000 > gpurun_out/r02_bench_line_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_line_n2.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'],'frac',d['roofline']['frac'],d['clocks']['sm_mhz'],d['verified'] and d['verified']['scan_kernel_ids_identical'])
ix=d.get('index_cfg3'); print('index', {k:ix[k] for k in ix if k not in ('runs','note','workload','best')} if ix else None)
for r in (ix or {}).get('runs',[]): print('   ',{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k!='phase_ms_rank0'})
PY
