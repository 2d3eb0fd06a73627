#!/bin/bash
# round 2, GPU run 26 (2 GPUs): fused latency with the minima bound, hybrid N = 2 with its full log, flat tests
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu26.log
exec > $L 2>&1
echo "== latency cfg1"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_latency.py 2>&1 | tail -2 | cut -c1-330
echo "== flat / hybrid tests"
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_hybrid.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
echo "== hybrid N=2, 2M docs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_hybrid.py --docs 2000000 --steps 5 > gpurun_out/r02_hybrid_n2.txt 2>&1
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_hybrid_n2.txt | tail -25 | cut -c1-1500
echo "== hybrid N=1, 2M docs (fusion scratch)"
CUDA_VISIBLE_DEVICES=1 timeout 900 python tools/bench_hybrid.py --docs 2000000 --steps 5 2>&1 | tail -1 | cut -c1-900
