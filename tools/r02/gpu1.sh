#!/usr/bin/env bash
# round 2, GPU call 1: box probe, full gpu suite (new parity matrix), bench line with verification, GEMM knob A/B
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_gpu1.log) 2>&1
echo "== box"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; df -h /tmp | tail -1
python -c "import torch,os; print(torch.cuda.device_count()); import glob; print(glob.glob(os.path.dirname(torch.__file__)+'/../nvidia/nccl/lib/*'))"
echo "== suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_line_a.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_line_a.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),d['clocks'],d.get('verified'))
print('flat',d.get('flat_scan'))
PY
echo "== knobs (headline only, 30 steps)"
for v in default backoff32 backoff128 singlepoll both; do
  if [ $v = default ]; then lib=""; else lib="build_variants/libb200search_$v.so"; fi
  for kps in 1 2; do
    B200_LIB_PATH=$lib B200_GEMM_KPS=$kps timeout 300 python bench.py --steps 30 --warmup 3 --headline-only 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v kps=$kps','value',round(d['value']),'ms',round(d['ms_per_step'],3),'kern_ms',round(d['roofline']['launch_ms'],3),'frac',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
  done
done
