#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu5.log) 2>&1
timeout 600 python tools/r02/debug_gemm3.py 2>&1 | tail -12
