#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp4.log) 2>&1
echo "=== one multicast gemm test first"
timeout 200 python -m pytest tests/test_gpu_flat.py -m gpu -q -k "test_gemm_path_matches_oracle and 33333-768-1024-10-2" --timeout 120 -x 2>&1 | tail -5
echo "=== all gemm tests"
timeout 500 python -m pytest tests/test_gpu_flat.py tests/test_gpu_edges.py -m gpu -q -k "gemm" --timeout 120 2>&1 | tail -5
for mc in 1 0; do
echo "=== bench multicast=$mc"
B200_GEMM_MULTICAST=$mc timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['launch_ms'],3),'TF',round(d['roofline']['achieved'],1),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done
