#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu12.log) 2>&1
for c in 0 1 2 16; do B200_IVF_COOP=$c timeout 200 python tools/r02/debug_ivf2.py 2>&1 | tail -7; done
