#!/bin/bash
set -u
echo "== pytest gpu (fast subset)"; timeout 900 python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -5
echo "== default lib"
timeout 300 python scripts/quick_bench.py --kernels 4,6 XSB200_WINDOW=32 XSB200_WINDOW=56 2>&1 | tail -12
for v in u1b5 u2b5; do
  [ -f scripts/exp/libxsb200_$v.so ] || continue
  echo "== variant $v"
  XSB200_GPU_LIB=$PWD/scripts/exp/libxsb200_$v.so timeout 300 python scripts/quick_bench.py --kernels 4 XSB200_WINDOW=32 2>&1 | tail -1
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -k regex:"xs_|sort_" -s 13 -c 13 --csv --log-file gpurun_out/launches_k4.csv python scripts/quick_bench.py --kernels 4 --reps 1 XSB200_WINDOW=32 2>&1 | tail -1
