#!/usr/bin/env bash
# per-launch durations of this library's kernels inside the bench step (never a bench value: runs under ncu)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 140 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_topk|gemm3_topk|topk_merge|flat_scan|f32_to_bf16|pad_rows|row_norms|split_tf32" -c 200 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launch_list_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r01_launches_final.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split('(')[0][:60]; agg.setdefault(k,[]).append(float(r[vi].replace(',',''))/1e6)
for k,v in agg.items(): print(f"{k:62s} launches {len(v):3d}  mean {sum(v)/len(v):8.3f} ms  total {sum(v):8.2f} ms")
PY
