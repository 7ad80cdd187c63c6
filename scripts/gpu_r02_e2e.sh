#!/bin/bash
# host-sample pipeline: [parity subset,] then chunk schedules (CONFIGS = e2e_bench.py configurations)
set -u
if [ -n "${TESTS:-}" ]; then timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$TESTS" 2>&1 | tail -4; fi
python scripts/e2e_bench.py ${CONFIGS:-""} 2>&1 | tail -24
