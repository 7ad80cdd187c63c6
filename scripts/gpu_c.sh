#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 "" XSB200_PACK_SAMPLES=0 2>&1 | tail -2
python scripts/e2e_bench.py "" XSB200_PACK_SAMPLES=0 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -3
