#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp11.log) 2>&1
timeout 600 python -m pytest tests/test_gpu_index.py tests/test_gpu_flat.py tests/test_gpu_edges.py tests/test_gpu_bm25.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python tools/ivf_latency.py
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|topk_merge|refine|flat_scan|pad_rows|normalize|gemm3" --csv --log-file gpurun_out/ivf_launches.csv python tools/ivf_latency.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/ivf_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:][-60:]:
    print(r[ki][:60], r[gi], r[vi])
PY
