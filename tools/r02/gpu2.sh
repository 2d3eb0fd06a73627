#!/usr/bin/env bash
# round 2, GPU call 2: the new index engine -- tests, then moderate-scale benchmarks
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu2.log) 2>&1
echo "== index tests"
timeout 900 python -m pytest tests/test_gpu_index.py -m gpu -q --timeout 600 -x 2>&1 | tail -25
echo "== bench ivf 2M x 768 MSTG"
timeout 600 python tools/bench_ivf.py --rows 2000000 --dim 768 --centres 2000 --type MSTG --nlist 2048 --nq 256 --nprobe 1,2,4,8,32 2>&1 | tail -12
echo "== bench ivf 2M x 96 IVFPQ unit"
timeout 600 python tools/bench_ivf.py --rows 2000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 1024 --m 96 --nq 10000 --nprobe 8,32 --truth-queries 256 2>&1 | tail -8
