#!/usr/bin/env bash
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu11.log) 2>&1
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all ./tests/cuda/coop_merge_test 2>&1 | grep -v "^=========$" | head -60
