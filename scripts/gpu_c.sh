#!/bin/bash
set -u
python scripts/quick_bench.py --kernels 6 --reps 3 "" XSB200_FUSE_GATHER=2 XSB200_FUSE_GATHER=0 2>&1 | tail -3
