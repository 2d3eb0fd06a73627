#!/usr/bin/env bash
# One gpurun call = many checks, each under its own timeout, all output kept in gpurun_out/.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
echo "=== gemm tests first (new code)"
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q -k "gemm" --timeout 200 -x 2>&1 | tail -25
echo "=== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 $BENCH_ARGS 2>&1 | tail -3
echo "=== bench TS off"
B200_GEMM_TS=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
if [ "$1" == "ncu" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_topk|topk_merge|pad_rows|f32_to_bf16|row_norms|flat_scan|normalize_rows" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
tail -2 gpurun_out/ncu_bench1.log | cut -c1-300
echo "=== ncu full capture of gemm_topk"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 -f -o gpurun_out/gemm_topk python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
ls -la gpurun_out
fi
