#!/bin/bash
# round 2, GPU run 33: both row passes' code loads in flight (SQ8 + PQ), 48-page items for many-query lists; index tests
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu33.log
exec > $L 2>&1
echo "== index tests"
timeout 900 python -m pytest tests/test_gpu_index.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
echo "== IVFSQ 20M x 768"
timeout 900 python tools/bench_ivf.py --rows 20000000 --dim 768 --centres 4000 --type IVFSQ --nlist 4096 --keep-raw 0 --nq 10240 --nprobe 1,4 --truth-queries 64 --reps 3 2>&1 | grep '"search"' | cut -c1-330
echo "== cfg4 shape"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 8,32 --truth-queries 8 --reps 3 2>&1 | grep '"search"' | cut -c1-330
