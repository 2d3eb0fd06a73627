#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp13.log) 2>&1
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ivfpq_scan" -s 2 -c 1 -o gpurun_out/r01_ivfpq_scan -f python tools/ivf_latency.py 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
