#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu13.log) 2>&1
timeout 200 python tools/r02/debug_ivf2.py 2>&1 | tail -7
./tests/cpp/callsite_compile
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | cut -c1-300
echo "== latency cfg 1"
timeout 300 python tools/bench_latency.py 2>&1 | tail -3
echo "== 10M x 768 MSTG nlist 8192"
timeout 900 python tools/bench_ivf.py --rows 10000000 --dim 768 --centres 10000 --type MSTG --nlist 8192 --nq 256 --nprobe 1,2,4,8,16 2>&1 | tail -8 | cut -c1-600
