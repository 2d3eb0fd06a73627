#!/bin/bash
# round 2, GPU run 34: final check of the tree as committed -- full suite, smoke, the driver's bench command
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu34.log
exec > $L 2>&1
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default"
timeout 900 python bench.py > gpurun_out/r02_bench_line_final2.json 2> gpurun_out/r02_bench_final2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_line_final2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),'verified',d['verified']['scan_kernel_ids_identical'],d['verified']['cpu_oracle_ids_identical'])
        print('flat',[round(x['frac_of_hbm_peak'],3) for x in d['flat_scan']],'lat',d['latency_cfg1']['resident_call_us'])
        ix=d['index_cfg3']; print('index qps@0.95',round(ix['qps_at_recall_0.95']),[(r['nprobe'],round(r['qps']),round(r['recall_at_10'],3),round(r['frac_of_hbm_peak_rank0'],2)) for r in ix['runs']])
PY
tail -2 gpurun_out/r02_bench_final2.err
