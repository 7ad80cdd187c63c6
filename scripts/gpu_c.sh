#!/bin/bash
set -u
echo "== pytest gpu (fast subset)"; timeout 900 python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -12
echo "== default lib"
timeout 300 python scripts/quick_bench.py --kernels 4,5,6 XSB200_WINDOW=32 XSB200_WINDOW=40 XSB200_WINDOW=64 2>&1 | tail -12
for v in u2b5 u3b4 u4b3 u1b5; do
  [ -f scripts/exp/libxsb200_$v.so ] || continue
  echo "== variant $v"
  XSB200_GPU_LIB=$PWD/scripts/exp/libxsb200_$v.so timeout 300 python scripts/quick_bench.py --kernels 4 XSB200_WINDOW=40 2>&1 | tail -1
done
