#!/bin/bash
# round 2, GPU run 23: warp-cooperative inserts for k >= 17 (unit test, micro-benchmark, k sweep, suite), BM25 with small tables,
# coarse probe A/B at the cfg-4 shape
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu23.log
exec > $L 2>&1
echo "== unit tests"
timeout 300 ./tests/cuda/list_append_test | tail -12
echo "== list_perf"
timeout 300 ./tests/cuda/list_perf 17000 1 | cut -c1-50
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== GEMM k sweep"
for k in 10 16 17 30 64 100 256; do
  timeout 300 python bench.py --steps 10 --warmup 3 --headline-only --k $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('k=$k','value',round(d['value']),'ms',round(d['ms_per_step'],3),'kern_ms',round(d['roofline']['launch_ms'],3))"
done
echo "== bm25"
timeout 600 python tools/bench_aux.py bm25 2>&1 | tail -1 | cut -c150-900
echo "== cfg4 shape: IVFPQ 100M x 96, nprobe 8/32, coarse path A/B"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 8,32 --truth-queries 64 --reps 3 --extra "coarse_path=1;coarse_path=2" 2>&1 | cut -c1-600
