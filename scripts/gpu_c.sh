#!/bin/bash
set -u
timeout 180 python scripts/quick_bench.py --kernels 6 --reps 3 2>&1 | tail -1
KERNELS=6 timeout 180 bash scripts/gpu_variants.sh notma
timeout 900 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
