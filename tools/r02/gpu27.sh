#!/bin/bash
# round 2, GPU run 27 (2 GPUs): fused-scan phase stamps, hybrid N = 2, ncu of the IVFPQ scan at the cfg-4 shape
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu27.log
exec > $L 2>&1
echo "== fused scan phase stamps"
CUDA_VISIBLE_DEVICES=0 B200_FUSED_DEBUG_TS=1 timeout 300 python tools/bench_latency.py 2>&1 | grep "fused ts" | sed -n '3,8p'
echo "== hybrid N=2, 2M docs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_hybrid.py --docs 2000000 --steps 5 > gpurun_out/r02_hybrid_n2.txt 2>&1
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/r02_hybrid_n2.txt | tail -4 | cut -c1-1200
echo "== ncu: IVFPQ scan at the cfg-4 shape (GPU 0)"
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:ivf_gemm_topk --launch-skip 2 -c 1 -o gpurun_out/r02_ivfpq_cfg4 -f python tools/bench_ivf.py --rows 100000000 --dim 96 --unit --type IVFPQ --metric L2 --nlist 4096 --m 96 --keep-raw 0 --nq 10000 --nprobe 32 --truth-queries 1 --reps 3 > gpurun_out/r02_ivfpq_ncu.txt 2>&1
tail -3 gpurun_out/r02_ivfpq_ncu.txt | cut -c1-300
