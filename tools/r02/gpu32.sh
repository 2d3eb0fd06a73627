#!/bin/bash
# round 2, GPU run 32: where does the SQ8 decoding scan spend its time (ncu), chunk sweep at the cfg-4 shape after the decode work
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu32.log
exec > $L 2>&1
echo "== cfg4 shape: chunk sweep"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 32 --truth-queries 8 --reps 3 --extra "pages_per_chunk=8;pages_per_chunk=16;pages_per_chunk=32;pages_per_chunk=48" 2>&1 | grep '"search"' | cut -c1-330
echo "== ncu: SQ8 scan, 20M x 768"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ivf_gemm_topk --launch-skip 2 -c 1 -o gpurun_out/r02_ivfsq -f python tools/bench_ivf.py --rows 20000000 --dim 768 --centres 4000 --type IVFSQ --nlist 4096 --keep-raw 0 --nq 10240 --nprobe 4 --truth-queries 1 --reps 3 > gpurun_out/r02_ivfsq_ncu.txt 2>&1
tail -2 gpurun_out/r02_ivfsq_ncu.txt | cut -c1-200
