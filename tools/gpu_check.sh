#!/usr/bin/env bash
# One gpurun call = many checks, each under its own timeout, all output kept in gpurun_out/.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
echo "=== scan / merge / binary tests"
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -k "not gemm" --timeout 300 -x 2>&1 | tail -25
echo "=== gemm tests"
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -k "gemm" --timeout 200 2>&1 | tail -40
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5
