#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:xs_sorted -s 1 -c 1 -f -o gpurun_out/prof_sorted python scripts/quick_bench.py --kernels 6 --reps 1 2>&1 | tail -1
