#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 "" XSB200_BIN_BITS=16 2>&1 | tail -2
KERNELS=6 bash scripts/gpu_variants.sh r8b3 r4 r16
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -5
