#!/bin/bash
python scripts/quick_bench.py --kernels 6 --reps 3 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m "gpu" 2>&1 | tail -3
