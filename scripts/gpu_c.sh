#!/bin/bash
python scripts/quick_bench.py --method history --kernels 0 XSB200_SWEEP=1 XSB200_SWEEP=0 2>&1 | tail -2
