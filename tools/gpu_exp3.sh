#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp3.log) 2>&1
timeout 600 python -m pytest tests/test_gpu_flat.py tests/test_gpu_edges.py tests/test_gpu_index.py -m gpu -q --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'TF',round(d['roofline']['achieved'],1),d['clocks']); [print(f) for f in d['flat_scan']]"
