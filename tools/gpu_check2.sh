#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check2.log) 2>&1
echo "=== bm25 + hybrid tests"
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_hybrid.py -m gpu -q --timeout 300 2>&1 | tail -30
echo "=== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
echo "=== TS vs streaming at d=512"
for ts in 1 0; do B200_GEMM_TS=$ts timeout 300 python bench.py --dim 512 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TS=$ts ms_per_step',round(d['ms_per_step'],3),'TF',round(d['roofline']['achieved'],1),d['clocks'])"; done
