#!/bin/bash
# XL (116.5 GiB, generated on the device): where the dense-segment kernel starts to pay
set -u
for dm in 0 8 20 64; do
  echo "== XL 17M  XSB200_DENSE_MIN=$dm"
  XSB200_DENSE_MIN=$dm timeout 600 xsbench_b200/xsbench -s XL -m event -k 6 --device-init --reps 3 2>&1 | grep -E "Device time|Lookups/s|checksum|failed" | tr '\n' ' '; echo
done
for dm in 0 20 64; do
  echo "== XL 1e9  XSB200_DENSE_MIN=$dm"
  XSB200_DENSE_MIN=$dm timeout 600 xsbench_b200/xsbench -s XL -m event -k 6 -l 1000000000 --device-init --reps 2 2>&1 | grep -E "Device time|Lookups/s|checksum|failed" | tr '\n' ' '; echo
done
