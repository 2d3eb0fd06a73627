#!/bin/bash
# round 2, GPU run 35: two-level tournament lists (opt-in): unit test, k sweep with and without, parity tests with it on
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu35.log
exec > $L 2>&1
echo "== unit test"
timeout 200 ./tests/cuda/list_append_test | tail -4
echo "== k sweep: default / tournament from k = 17"
for mode in 0 17; do for k in 30 64 100; do
  B200_LIST_TOURN_MIN_K=$mode timeout 200 python bench.py --steps 10 --warmup 3 --headline-only --k $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tourn_min_k=$mode k=$k','value',round(d['value']),'ms',round(d['ms_per_step'],3))"
done; done
echo "== parity with the tournament form on (GEMM variants at 2M rows, index tests)"
B200_LIST_TOURN_MIN_K=17 timeout 600 python -m pytest tests/test_gpu_gemm_scale.py tests/test_gpu_index.py -m gpu -x -q 2>&1 | tail -3
