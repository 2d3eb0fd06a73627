#!/bin/bash
# round 2, GPU run 31 (8 GPUs): the driver's N = 8 launch of bench.py (shorter index extra) and the sharded hybrid step
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu31.log
exec > $L 2>&1
nvidia-smi -L | wc -l
echo "== bench N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 3 --index-rows 32000000 > gpurun_out/r02_bench_line_n8.json 2> gpurun_out/r02_bench_n8.err
grep '^{' gpurun_out/r02_bench_line_n8.json | tail -1 | cut -c1-3500
tail -3 gpurun_out/r02_bench_n8.err | cut -c1-300
echo "== hybrid N=8, 4M docs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/bench_hybrid.py --docs 4000000 --steps 5 > gpurun_out/r02_hybrid_n8.txt 2>&1
grep '^{' gpurun_out/r02_hybrid_n8.txt | tail -1 | cut -c1-1200
grep -i "error\|Traceback" gpurun_out/r02_hybrid_n8.txt | head -5
