#!/bin/bash
# one-kernel capture: KERNEL=regex SKIP=n OUT=name KID=kernel-id
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-xs_sorted} -s ${SKIP:-1} -c 1 -f -o gpurun_out/${OUT:-prof} python scripts/quick_bench.py --kernels ${KID:-6} --reps 1 2>&1 | tail -1
