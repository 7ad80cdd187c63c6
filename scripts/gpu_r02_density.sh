#!/bin/bash
# -k 6 phase times against the lookup count (density of the sorted lookups): what a chunk of a host-sample call costs
set -u
for n in ${NS:-17000000 8500000 5666666 4250000 2125000}; do
  python scripts/quick_bench.py --kernels 6 --reps ${REPS:-5} --lookups $n ${CONFIGS:-""} 2>&1 | tail -${TAIL:-1} | sed "s/^/n=$n /"
done
