#!/bin/bash
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -3
python scripts/quick_bench.py --grid nuclide --kernels 0,4,6 XSB200_NUCLIDE_BUCKETS=1 XSB200_NUCLIDE_BUCKETS=0 2>&1 | tail -6
python scripts/quick_bench.py --grid nuclide --method history --kernels 0 2>&1 | tail -1
