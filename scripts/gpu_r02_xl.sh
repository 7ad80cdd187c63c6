#!/bin/bash
# XL (116.5 GiB, built on the device) and 10^9-lookup runs through the C driver; sort-key width experiment
set -u
echo "== key width (lookup phase penalty of coarser energy keys; still 3 passes)"
python scripts/quick_bench.py --kernels 6 --reps 4 "" XSB200_KEY_LO_BIT=10 XSB200_KEY_LO_BIT=12 XSB200_KEY_LO_BIT=14 2>&1 | tail -4
echo "== large, 10^9 lookups"
timeout 600 xsbench_b200/xsbench -s large -m event -k 6 -l 1000000000 --device-init --reps 2 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|failed" | tr '\n' ' '; echo
echo "== XL via C driver (device-side generation)"
for k in 6 0; do
  timeout 600 xsbench_b200/xsbench -s XL -m event -k $k --device-init --reps 3 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|failed" | tr '\n' ' '; echo
done
timeout 600 xsbench_b200/xsbench -s XL -m event -k 6 -l 1000000000 --device-init --reps 2 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum|failed" | tr '\n' ' '; echo
