#!/bin/bash
# per-material lookup-phase times + one ncu capture of xs_dense_kernel on the non-fuel samples only
set -u
mkdir -p gpurun_out
timeout 600 python scripts/exp/split_by_material.py 2>&1 | tail -16 | tee gpurun_out/r02_split_by_material.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xs_dense -s 1 -c 1 -f -o gpurun_out/r02_dense_nonfuel python scripts/exp/nonfuel_only.py 2>&1 | tail -3
