#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu15.log) 2>&1
echo "== bench N=1 (headline + extras incl. 100M index)"
timeout 1500 python bench.py --steps 50 --warmup 3 > gpurun_out/r02_bench_line_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_line_b.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])
ix=d.get('index_cfg3'); print('index', {k:ix[k] for k in ix if k not in ('runs','note','workload','best')} if ix else None)
for r in (ix or {}).get('runs',[]): print('   ',{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()})
PY
