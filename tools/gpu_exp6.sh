#!/usr/bin/env bash
# cluster-of-8 multicast (4 pairs share a corpus tile) vs cluster-of-4; flat scan cold vs after GEMM
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_exp6.log) 2>&1
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1','ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['launch_ms'],3),'TF',round(d['roofline']['achieved'],1),'e2e',round(d['e2e']['value']),d['clocks']['sm_mhz'],d['clocks']['reasons'], [round(x['GB_per_s']) for x in d.get('flat_scan',[])])"; }
for mc in 4 2 4 2; do
B200_GEMM_MULTICAST=$mc timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | line mc$mc
done
for slack in 0 2 8; do
B200_GEMM_SYNC_SLACK=$slack timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | line mc4_slack$slack
done
