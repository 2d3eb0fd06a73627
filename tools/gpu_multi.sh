#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_multi.log) 2>&1
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
for n in 1 $N; do
echo "=== bench --gpus $n"
if [ $n == 1 ]; then
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1600
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 2>&1 | tail -3 | cut -c1-1600
fi
done
