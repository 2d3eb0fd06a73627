#!/usr/bin/env bash
mkdir -p gpurun_out
N=${1:-2}
exec > >(tee gpurun_out/gpu_multi_$N.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in $N; do
echo "=== bench --gpus $n"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 2>&1 | tail -2 | cut -c1-2500
done
