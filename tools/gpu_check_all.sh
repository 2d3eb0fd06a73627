#!/usr/bin/env bash
# full GPU validation: suite, smoke, BM25 aux bench, short bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check_all.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_aux.py bm25
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms_per_step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
