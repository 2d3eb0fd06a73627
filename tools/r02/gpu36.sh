#!/bin/bash
# round 2, GPU run 36 (last seconds of the budget): headline with constant list stride, tournament default on, GEMM parity at 2M rows
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu36.log
exec > $L 2>&1
timeout 60 python bench.py --steps 10 --warmup 3 --headline-only --k 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('k=10 value',round(d['value']),'ms',round(d['ms_per_step'],3))"
timeout 120 python -m pytest tests/test_gpu_gemm_scale.py -m gpu -x -q 2>&1 | tail -2
