#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (fast subset)"; timeout 900 python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -4
export XSB200_GPU_LIB=$PWD/scripts/exp/libxsb200_u2b5.so
python scripts/quick_bench.py --kernels 4,6 XSB200_WINDOW=40 XSB200_WINDOW=400 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 25 -c 1 -f -o gpurun_out/prof_window python scripts/quick_bench.py --kernels 4 --reps 1 XSB200_WINDOW=40 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"xs_|sort_" -s 22 -c 22 --csv --log-file gpurun_out/launches_k4.csv python scripts/quick_bench.py --kernels 4 --reps 1 XSB200_WINDOW=40 2>&1 | tail -1
