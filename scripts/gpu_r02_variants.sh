#!/bin/bash
# usage: VARIANTS="main a b" [TESTS=expr] [CONFIGS="default X=1"] gpu_r02_variants.sh
set -u
if [ -n "${TESTS:-}" ]; then timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$TESTS" 2>&1 | tail -4; fi
for v in ${VARIANTS:-main}; do
  lib=$PWD/xsbench_b200/variants/libxsb200_$v.so
  [ "$v" = main ] && lib=$PWD/xsbench_b200/libxsb200.so
  cfgs=()
  for c in ${CONFIGS:-default}; do if [ "$c" = default ]; then cfgs+=(""); else cfgs+=("$c"); fi; done
  echo "=== $v"
  XSB200_GPU_LIB=$lib timeout 600 python scripts/quick_bench.py --kernels ${KIDS:-6} --reps ${REPS:-5} --grid ${GRID:-unionized} --size ${SIZE:-large} "${cfgs[@]}" 2>&1 | tail -${NCONF:-4}
done
