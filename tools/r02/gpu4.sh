#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu4.log) 2>&1
echo "== 400k x 768, nlist 512 (small k-means path)"
timeout 600 python tools/bench_ivf.py --rows 400000 --dim 768 --centres 400 --type MSTG --nlist 512 --nq 256 --nprobe 1,8 2>&1 | tail -6
echo "== 2M x 768, nlist 2048 (tensor-core k-means path)"
timeout 600 python tools/bench_ivf.py --rows 2000000 --dim 768 --centres 2000 --type MSTG --nlist 2048 --nq 256 --nprobe 1,8,32 2>&1 | tail -8
