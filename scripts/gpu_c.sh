#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -5
python scripts/quick_bench.py --kernels 4,6 --reps 3 "" XSB200_KEY_LO_BIT=16 XSB200_KEY_LO_BIT=20 XSB200_KEY_LO_BIT=24 2>&1 | tail -8
# k6 fuel window: count launches first
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 14 -c 1 -f -o gpurun_out/prof_window_k6 python scripts/quick_bench.py --kernels 6 --reps 1 2>&1 | tail -1
