#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp8.log) 2>&1
nproc
timeout 300 python tools/bench_aux.py ingest
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line.json; cat gpurun_out/bench_line.json
