#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu14.log) 2>&1
echo "== flat tests"
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_edges.py tests/test_gpu_properties.py -m gpu -q --timeout 600 2>&1 | tail -4
echo "== latency"
timeout 300 python tools/bench_latency.py 2>&1 | tail -2
echo "== bench N=1 (headline + extras incl. 100M index)"
timeout 1500 python bench.py --steps 50 --warmup 3 > gpurun_out/r02_bench_line_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_line_b.json') if l.startswith('{')][-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])
print('verified',d.get('verified') and {k:d['verified'][k] for k in ('scan_kernel_ids_identical','cpu_oracle_ids_identical')})
print('flat',[(f.get('queries_per_pass'),round(f.get('GB_per_s',0)),round(f.get('frac_of_hbm_peak',0),3)) for f in d['flat_scan']])
print('latency',d.get('latency_cfg1'))
ix=d.get('index_cfg3'); print('index', {k:ix[k] for k in ix if k not in ('runs','note','workload')} if ix else None)
for r in (ix or {}).get('runs',[]): print('   ',{k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()})
print('cpu',d.get('cpu_baseline'))
PY
