#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu18.log) 2>&1
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | cut -c1-300
echo "== bm25 bench DAAT / TAAT"
timeout 600 python tools/bench_aux.py bm25 2>&1 | tail -1 | cut -c150-900
B200_BM25_TAAT=1 timeout 600 python tools/bench_aux.py bm25 2>&1 | tail -1 | cut -c150-900
echo "== GEMM k sweep (lists in smem only with >= N stages)"
for ms in 5 2; do for k in 30 64 100; do
  B200_GEMM_LIST_SMEM_MIN_STAGES=$ms timeout 300 python bench.py --steps 10 --warmup 3 --headline-only --k $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('min_stages=$ms k=$k','value',round(d['value']),'ms',round(d['ms_per_step'],3),'kern_ms',round(d['roofline']['launch_ms'],3))"
done; done
echo "== cfg4 shape: IVFPQ 100M x 96, nlist 4096, m 96, nq 10000"
timeout 1500 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 8,32 --truth-queries 64 --reps 3 2>&1 | tail -6 | cut -c1-600
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --headline-only > /dev/null 2>&1; wc -l gpurun_out/r02_launches_bench.csv
echo "== ncu full: ivf scan v2 (10M x 768, nprobe 1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_gemm_topk_kernel -s 3 -c 1 -o gpurun_out/r02_ivf_scan_v2 -f python tools/bench_ivf.py --rows 10000000 --dim 768 --centres 10000 --type MSTG --nlist 8192 --keep-raw 0 --nq 256 --nprobe 1 --reps 3 --truth-queries 16 2>&1 | tail -2 | cut -c1-300
