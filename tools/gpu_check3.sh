#!/usr/bin/env bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check3.log) 2>&1
echo "=== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_line.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}, d['e2e']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'])
print('flat_scan', d.get('flat_scan'))
print('cpu', d.get('cpu_baseline'))
PY
echo "=== ncu flat_scan"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_scan -c 2 -f -o gpurun_out/flat_scan python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
tail -2 gpurun_out/ncu_bench3.log | cut -c1-200
echo "=== ncu launch list (library kernels)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_topk|topk_merge|pad_rows|f32_to_bf16|row_norms|flat_scan|normalize_rows" -c 60 --csv --log-file gpurun_out/launches_b200.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
tail -1 gpurun_out/ncu_bench1.log | cut -c1-200
