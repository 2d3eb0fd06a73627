#!/bin/bash
# round 2, GPU run 30: full suite + the driver's bench command + smoke, after the decode / tail changes
cd "$(dirname "$0")/../.."
L=gpurun_out/r02_gpu30.log
exec > $L 2>&1
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== IVFSQ (SQ8 decoder without I2F) 20M x 768"
timeout 900 python tools/bench_ivf.py --rows 20000000 --dim 768 --centres 4000 --type IVFSQ --nlist 4096 --keep-raw 0 --nq 10240 --nprobe 1,4 --truth-queries 64 --reps 3 2>&1 | grep '"search"' | cut -c1-330
echo "== latency"
timeout 300 python tools/bench_latency.py 2>&1 | tail -2 | cut -c1-200
echo "== bench default"
timeout 1500 python bench.py > gpurun_out/r02_bench_line_final.json 2> gpurun_out/r02_bench_final.err
tail -c 6000 gpurun_out/r02_bench_line_final.json
tail -3 gpurun_out/r02_bench_final.err
echo "== bench reference arm (short)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
