#!/bin/bash
python scripts/quick_bench.py --kernels 4 2>&1 | tail -1
for v in NOMATH NOSHFL NOLDG NOINDEX; do echo $v; XSB200_GPU_LIB=$PWD/scripts/exp/libxsb200_$v.so python scripts/quick_bench.py --kernels 4 2>&1 | tail -1; done
