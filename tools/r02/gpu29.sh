#!/bin/bash
# round 2, GPU run 29: fused tail fixed (aligned minima), L2 prefetch of the next page in the decoding scan
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu29.log
exec > $L 2>&1
echo "== unit tests"
timeout 300 ./tests/cuda/list_append_test | tail -2
timeout 300 ./tests/cuda/coop_merge_test | tail -1
echo "== latency cfg1"
B200_FUSED_DEBUG_TS=1 timeout 300 python tools/bench_latency.py 2> gpurun_out/r02_fused_ts.txt | tail -2 | cut -c1-330
sed -n '3,5p' gpurun_out/r02_fused_ts.txt | cut -c1-300
echo "== index / flat / property tests"
timeout 1500 python -m pytest tests/test_gpu_index.py tests/test_gpu_flat.py tests/test_gpu_properties.py tests/test_gpu_hybrid.py -m gpu -x -q 2>&1 | tail -4
echo "== cfg4 shape: IVFPQ 100M x 96, nprobe 8/32"
timeout 1200 python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 8,32 --truth-queries 64 --reps 3 2>&1 | grep '"search"' | cut -c1-600
echo "== IVFSQ (SQ8 decoder) 20M x 768, nprobe 4"
timeout 900 python tools/bench_ivf.py --rows 20000000 --dim 768 --centres 4000 --type IVFSQ --nlist 4096 --keep-raw 0 --nq 10240 --nprobe 1,4 --truth-queries 64 --reps 3 2>&1 | grep '"search"' | cut -c1-500
