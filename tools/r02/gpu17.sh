#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu17.log) 2>&1
echo "== bm25 / hybrid / property tests (DAAT kernel)"
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_hybrid.py tests/test_gpu_properties.py tests/test_gpu_index.py -m gpu -q --timeout 600 2>&1 | tail -6 | cut -c1-300
echo "== bm25 bench DAAT"
timeout 600 python tools/bench_aux.py bm25 2>&1 | tail -1 | cut -c1-600
echo "== bm25 bench TAAT (round 1 kernel)"
B200_BM25_TAAT=1 timeout 600 python tools/bench_aux.py bm25 2>&1 | tail -1 | cut -c1-600
echo "== 100M x 768 MSTG, chunking sweep"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 768 --centres 10000 --type MSTG --nlist 16384 --keep-raw 0 --nq 256 --nprobe 1,2,4 --truth-queries 128 --extra ";pages_per_chunk=4;pages_per_chunk=8;pages_per_chunk=16" 2>&1 | tail -16 | cut -c1-420
