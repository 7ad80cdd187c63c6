#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 2>&1 | tail -1
KERNELS=6 bash scripts/gpu_variants.sh p4c16 p2c16
