#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu9.log) 2>&1
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
echo "== latency cfg 1"
timeout 300 python tools/bench_latency.py 2>&1 | tail -3
echo "== 10M x 768 MSTG nlist 8192"
timeout 900 python tools/bench_ivf.py --rows 10000000 --dim 768 --centres 10000 --type MSTG --nlist 8192 --nq 256 --nprobe 1,2,4,8,16 2>&1 | tail -9 | cut -c1-600
echo "== 5M x 96 IVFPQ"
timeout 900 python tools/bench_ivf.py --rows 5000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --nq 10000 --nprobe 8,32 --truth-queries 256 2>&1 | tail -5 | cut -c1-600
