#!/bin/bash
# one-kernel ncu --set full capture: KERNEL=regex SKIP=n OUT=name KID=kernel-id [ENVS="A=1 B=2"]
set -u
mkdir -p gpurun_out
for kv in ${ENVS:-}; do export "$kv"; done
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-xs_dense} -s ${SKIP:-1} -c 1 -f -o gpurun_out/${OUT:-prof} python scripts/quick_bench.py --kernels ${KID:-6} --reps 1 2>&1 | tail -2
