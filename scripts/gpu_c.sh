#!/bin/bash
set -u
timeout 1200 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_now.json; cut -c1-1500 gpurun_out/bench_now.json
