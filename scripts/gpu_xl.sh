#!/bin/bash
set -u
timeout 1500 python -m pytest tests -x -q -m "gpu and slow" 2>&1 | tail -5
echo "== XL timing via C driver (device-side generation)"
for k in 6 4; do
  timeout 600 xsbench_b200/xsbench -s XL -m event -k $k --device-init --reps 3 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|failed" | tr '\n' ' '; echo
done
XSB200_SORTED_KERNEL=0 timeout 600 xsbench_b200/xsbench -s XL -m event -k 6 --device-init --reps 3 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|failed" | tr '\n' ' '; echo
