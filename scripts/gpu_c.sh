#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 2>&1 | tail -1
python scripts/exp/split_by_material.py 2>&1 | head -3
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
