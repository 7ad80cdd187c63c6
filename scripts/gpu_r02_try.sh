#!/bin/bash
# quick iteration: [TESTS=pytest -k expr] then quick_bench over CONFIGS (space separated ENV=V[,ENV=V] strings; "default" = none)
set -u
mkdir -p gpurun_out
if [ -n "${TESTS:-}" ]; then timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$TESTS" 2>&1 | tail -4; fi
cfgs=()
for c in ${CONFIGS:-default}; do if [ "$c" = default ]; then cfgs+=(""); else cfgs+=("$c"); fi; done
timeout 900 python scripts/quick_bench.py --kernels ${KIDS:-6} --reps ${REPS:-5} --grid ${GRID:-unionized} --size ${SIZE:-large} "${cfgs[@]}" 2>&1 | tail -20
