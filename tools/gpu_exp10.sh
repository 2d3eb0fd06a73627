#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp10.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 600 python tools/bench_aux.py mstg bm25 flat10k
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),[round(x['GB_per_s']) for x in d['flat_scan']], d['fp32_batch']['ms_per_launch'])"
