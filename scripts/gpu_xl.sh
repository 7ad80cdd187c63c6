#!/bin/bash
set -u
free -g | head -2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_pass or xl_problem" 2>&1 | tail -5
echo "== XL timing via C driver"
timeout 900 xsbench_b200/xsbench -s XL -m event -k 4 --reps 3 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|Allocated|failed"
