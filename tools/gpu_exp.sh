#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp.log) 2>&1
for dbg in 0 1; do
  echo "=== B200_GEMM_DEBUG=$dbg"
  B200_GEMM_DEBUG=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'clock',d['clocks'],'kernel_ms',round(d['roofline']['launch_ms'],3))
except Exception as e: print('ERR',e)
"
done
echo "=== streaming pairs (B200_GEMM_TS=0)"
B200_GEMM_TS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
