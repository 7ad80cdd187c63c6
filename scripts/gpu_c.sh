#!/bin/bash
set -u
for l in 17000000 1000000000; do
XSB200_BANDS=2 XSB200_BAND_INDEX=0 timeout 600 xsbench_b200/xsbench -s large -m event -l $l -k 6 --device-init 2>&1 | grep -E "Device time|Phases|Lookups/s" | tr '\n' ' '; echo
XSB200_BANDS=2 XSB200_BAND_INDEX=1 timeout 600 xsbench_b200/xsbench -s large -m event -l $l -k 6 --device-init 2>&1 | grep -E "Device time|Phases|Lookups/s" | tr '\n' ' '; echo
done
