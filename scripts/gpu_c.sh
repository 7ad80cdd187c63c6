#!/bin/bash
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -3
python scripts/quick_bench.py --kernels 0,1 2>&1 | tail -2
python scripts/quick_bench.py --grid hash --kernels 0 2>&1 | tail -1
python scripts/quick_bench.py --grid nuclide --kernels 0 2>&1 | tail -1
