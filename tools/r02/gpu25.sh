#!/bin/bash
# round 2, GPU run 25 (2 GPUs): multi-rank tests, hybrid tool at N = 1 and N = 2 (same table: checksums must agree), fused latency
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu25.log
exec > $L 2>&1
nvidia-smi -L
echo "== latency cfg1 (staged tail)"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_latency.py 2>&1 | tail -2 | cut -c1-400 &
echo "== hybrid N=1 (GPU 1), 2M docs"
CUDA_VISIBLE_DEVICES=1 timeout 900 python tools/bench_hybrid.py --docs 2000000 --steps 5 2>&1 | tail -1
wait
echo "== hybrid N=2, 2M docs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_hybrid.py --docs 2000000 --steps 5 2>&1 | tail -1
echo "== multi-rank tests + flat/fused tests"
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_flat.py -m gpu -x -q 2>&1 | tail -5
