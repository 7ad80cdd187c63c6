#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 14 -c 1 -f -o gpurun_out/prof_window python scripts/quick_bench.py --kernels 4 --reps 1 XSB200_WINDOW=32 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"xs_|sort_" -s 14 -c 14 --csv --log-file gpurun_out/launches_k4.csv python scripts/quick_bench.py --kernels 4 --reps 1 XSB200_WINDOW=32 2>&1 | tail -1
