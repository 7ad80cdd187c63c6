#!/bin/bash
python scripts/quick_bench.py --kernels 6,4,0 --reps 3 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
