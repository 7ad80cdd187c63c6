#!/bin/bash
# A/B the libxsb200 variants built by scripts/build_variants.sh
set -u
for v in "$@"; do
  echo "=== $v"
  XSB200_GPU_LIB=$PWD/xsbench_b200/variants/libxsb200_$v.so python scripts/quick_bench.py --kernels ${KERNELS:-4,6} --reps 3 2>&1 | tail -2
done
