#!/bin/bash
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -5
python -m pytest tests -m "gpu" -x -q -k "official_large" 2>&1 | tail -3
xsbench_b200/xsbench -m event -s large -k 4 --device-init --reps 3 2>&1 | grep -E "Building|Allocated|Device time|Lookups/s|checksum"
