#!/bin/bash
# round 2, GPU run 24: dedicated coarse probe (suite + cfg-4 shape), ncu of the fused single-query scan, hybrid tool
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu24.log
exec > $L 2>&1
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== cfg4 shape: IVFPQ 100M x 96, nprobe 8/32/64: coarse select"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 8,32,64 --truth-queries 64 --reps 3 --extra "coarse_path=3" 2>&1 | cut -c1-600
echo "== hybrid (1 GPU, 1M docs)"
timeout 900 python tools/bench_hybrid.py --docs 1000000 --steps 5 2>&1 | tail -1
echo "== ncu: fused single-query scan"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flat_scan_kernel --launch-skip 50 -c 1 -o gpurun_out/r02_fused_scan -f python tools/bench_latency.py > gpurun_out/r02_fused_ncu.txt 2>&1
tail -3 gpurun_out/r02_fused_ncu.txt
