#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp2.log) 2>&1
echo "=== gemm tests"
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -k "gemm" --timeout 200 -x 2>&1 | tail -5
for dbg in 0 1; do for rows in 1250000 10000000; do
  B200_GEMM_DEBUG=$dbg timeout 300 python bench.py --rows $rows --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('debug',$dbg,'rows',$rows,'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['launch_ms'],3),'TF',round(d['roofline']['achieved'],1),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done; done
