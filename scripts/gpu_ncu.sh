#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 17 -c 1 -f -o gpurun_out/prof_window python scripts/quick_bench.py --kernels 4 --reps 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 21 -c 1 -f -o gpurun_out/prof_small python scripts/quick_bench.py --kernels 4 --reps 1 2>&1 | tail -1
