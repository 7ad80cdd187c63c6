#!/bin/bash
# usage: gpu_variants_dense.sh VARIANT...   ("main" = the in-tree library)
set -u
for v in "$@"; do
  echo "=== $v"
  lib=$PWD/xsbench_b200/variants/libxsb200_$v.so
  [ "$v" = main ] && lib=$PWD/xsbench_b200/libxsb200.so
  XSB200_GPU_LIB=$lib python scripts/quick_bench.py --kernels 6 --reps 4 ${CONFIGS:-XSB200_DENSE_MIN=64} 2>&1 | tail -${NCONF:-1}
done
