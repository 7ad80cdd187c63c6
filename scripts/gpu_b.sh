#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (fast subset)"; timeout 900 python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -4
echo "== quick bench"
timeout 600 python scripts/quick_bench.py --kernels 0,4,5,6 XSB200_GATHER=0 XSB200_SWEEP=0 2>&1 | tail -20
ncu --set full --clock-control none --import-source on -k regex:xs_sweep_kernel -s 12 -c 1 -f -o gpurun_out/prof_sweep python scripts/quick_bench.py --kernels 4 --reps 1 2>&1 | tail -1
