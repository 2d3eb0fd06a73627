#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu8.log) 2>&1
echo "== ncu: ivf_gemm_topk_kernel, 2M x 768 MSTG nprobe 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_gemm_topk_kernel -s 3 -c 1 -o gpurun_out/r02_ivf_scan_v1 -f \
  python tools/bench_ivf.py --rows 2000000 --dim 768 --centres 2000 --type MSTG --nlist 2048 --nq 256 --nprobe 1 --reps 3 2>&1 | tail -4 | cut -c1-400
ls -la gpurun_out/*.ncu-rep
