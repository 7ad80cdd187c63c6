#!/bin/bash
# lookups per lane of the dense kernel vs density: -l LOOKUPS sweeps x DENSE_MIN x variants
set -u
for L in ${LOOKUPS:-1000000 2000000 4000000 8000000}; do
  for v in ${VARIANTS:-main pl2 pl1}; do
    lib=$PWD/xsbench_b200/variants/libxsb200_$v.so
    [ "$v" = main ] && lib=$PWD/xsbench_b200/libxsb200.so
    echo "=== lookups $L  $v"
    XSB200_GPU_LIB=$lib timeout 600 python scripts/quick_bench.py --kernels 6 --reps 3 --lookups $L ${CONFIGS:-"" XSB200_DENSE_MIN=4} 2>&1 | tail -${NCONF:-2}
  done
done
