#!/bin/bash
# round 2, GPU run 22: shared per-query bound in the IVF scan (A/B), GEMM k sweep back on rescan lists, hybrid tool
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu22.log
exec > $L 2>&1
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== GEMM k sweep (rescan lists in smem)"
for k in 10 30 100; do
  timeout 300 python bench.py --steps 10 --warmup 3 --headline-only --k $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('k=$k','value',round(d['value']),'ms',round(d['ms_per_step'],3),'kern_ms',round(d['roofline']['launch_ms'],3))"
done
echo "== cfg4 shape: IVFPQ 100M x 96, nprobe 32: bound on/off x chunk"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 32 --truth-queries 64 --reps 3 --extra "shared_bound=0;shared_bound=1;shared_bound=1,pages_per_chunk=96;shared_bound=0,pages_per_chunk=96" 2>&1 | cut -c1-600
echo "== 100M x 768 MSTG (SQ8 codes), nq 10240, nprobe 1/2/4: bound on/off"
timeout 1500 python tools/bench_ivf.py --rows 100000000 --dim 768 --centres 10000 --type MSTG --nlist 16384 --keep-raw 0 --nq 10240 --nprobe 1,2,4 --truth-queries 64 --reps 3 --extra "shared_bound=1;shared_bound=0" 2>&1 | cut -c1-600
echo "== hybrid (1 GPU, 1M docs)"
timeout 900 python tools/bench_hybrid.py --docs 1000000 --steps 5 2>&1 | tail -2
