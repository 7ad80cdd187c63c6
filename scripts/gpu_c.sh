#!/bin/bash
S=$(date +%s.%N)
xsbench_b200/xsbench -m event -s XL -k 6 --device-init --reps 3 2>&1 | grep -E "Building|Allocated|Device time|Phases|Lookups/s|checksum|failed"
E=$(date +%s.%N); echo "wall: $(echo "$E - $S" | bc) s"
S=$(date +%s.%N)
xsbench_b200/xsbench -m event -s XL -l 1000000 -k 4 --device-init 2>&1 | grep -E "checksum|failed"
E=$(date +%s.%N); echo "wall: $(echo "$E - $S" | bc) s"
