#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== full gpu test suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
