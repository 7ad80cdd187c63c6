#!/bin/bash
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -3
python scripts/quick_bench.py --kernels 4,6 2>&1 | tail -2
python scripts/quick_bench.py --method history --kernels 0 2>&1 | tail -1
