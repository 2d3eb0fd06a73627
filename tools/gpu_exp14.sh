#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp14.log) 2>&1
REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,smsp__inst_executed.sum --clock-control none -k regex:"ivfpq_scan|topk_merge|refine|flat_scan|pad_rows|normalize|gemm3" --csv --log-file gpurun_out/ivf_launches2.csv python tools/ivf_latency.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ivf_launches2.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); mi=hdr.index('Metric Name'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[ii],{'k':r[ki][:40],'g':r[gi]})[r[mi]]=r[vi]
for i,(k,v) in enumerate(cur.items()):
    if i<24: print(v)
PY
