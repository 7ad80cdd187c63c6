#!/bin/bash
set -u
for v in p4c16 p4c8; do echo "== $v"; XSB200_GPU_LIB=$PWD/xsbench_b200/variants/libxsb200_$v.so python scripts/exp/split_by_material.py 2>&1 | head -4; done
