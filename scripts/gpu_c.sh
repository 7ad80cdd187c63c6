#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6,4,0 --reps 3 "" XSB200_INDEX_COLUMNS=0 2>&1 | tail -6
python scripts/quick_bench.py --method history --kernels 0 --reps 2 "" XSB200_INDEX_COLUMNS=0 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
