#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp12.log) 2>&1
timeout 600 python -m pytest tests/test_gpu_index.py tests/test_gpu_hybrid.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python tools/ivf_latency.py
timeout 600 python tools/bench_aux.py mstg ivfpq
