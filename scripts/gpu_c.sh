#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 2>&1 | tail -1
KERNELS=6 bash scripts/gpu_variants.sh sp2 sp8
python scripts/exp/split_by_material.py 2>&1 | tail -15
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
