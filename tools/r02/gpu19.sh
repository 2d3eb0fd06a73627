#!/bin/bash
# round 2, GPU run 19: append-mode lists (unit test, GEMM k sweep, IVF per-thread items), fused-scan latency, hybrid tool
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu19.log
exec > $L 2>&1
echo "== unit tests"
timeout 300 ./tests/cuda/list_append_test | tail -12
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== GEMM k sweep (append lists; min_stages 4 = default, 2 = always smem when they fit, 7 = always global)"
for ms in 4 2 7; do for k in 10 30 64 100; do
  B200_GEMM_LIST_SMEM_MIN_STAGES=$ms timeout 300 python bench.py --steps 10 --warmup 3 --headline-only --k $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('min_stages=$ms k=$k','value',round(d['value']),'ms',round(d['ms_per_step'],3),'kern_ms',round(d['roofline']['launch_ms'],3))"
done; done
echo "== cfg4 shape: IVFPQ 100M x 96 (append lists), chunk sweep"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 32 --truth-queries 64 --reps 3 --extra "pages_per_chunk=16;pages_per_chunk=32;pages_per_chunk=96" 2>&1 | cut -c1-700
echo "== 100M x 768 MSTG, nprobe 1/2/4, chunk sweep"
timeout 1500 python tools/bench_ivf.py --rows 100000000 --dim 768 --type MSTG --metric COSINE --nq 10240 --nprobe 1,2,4 --truth-queries 64 --reps 3 --extra "pages_per_chunk=16;pages_per_chunk=32;pages_per_chunk=64" 2>&1 | cut -c1-700
echo "== latency cfg1"
timeout 300 python tools/bench_latency.py 2>&1 | tail -12
echo "== hybrid (1 GPU, 2M docs)"
timeout 900 python tools/bench_hybrid.py --docs 2000000 --steps 5 2>&1 | tail -3
