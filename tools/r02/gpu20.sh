#!/bin/bash
# round 2, GPU run 20/21: rescan vs append lists without counters, with the tensor-core kernels' shared-memory footprint (tiny L1)
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu21.log
exec > $L 2>&1
echo "== list_perf (no counters, 220 KB smem carve-out)"
timeout 600 ./tests/cuda/list_perf 17000 1
