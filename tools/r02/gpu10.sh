#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu10.log) 2>&1
./tests/cpp/callsite_compile
timeout 600 python -m pytest tests/test_gpu_index.py -m gpu -q --timeout 600 2>&1 | grep -v "^$" | tail -80 | cut -c1-400
timeout 300 python tools/r02/debug_ivf.py 2>&1 | tail -32 | cut -c1-300
