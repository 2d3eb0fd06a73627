#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp5.log) 2>&1
timeout 500 python -m pytest tests/test_gpu_flat.py tests/test_gpu_edges.py -m gpu -q -k "gemm" --timeout 120 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for k in 30 100 128; do
timeout 300 python bench.py --k $k --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('k',$k,'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['launch_ms'],3),'TF',round(d['roofline']['achieved'],1),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done
