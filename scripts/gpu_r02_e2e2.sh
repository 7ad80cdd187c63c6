#!/bin/bash
# e2e A/B (environment only): dense threshold and arithmetic at the chunk densities of the host-sample call
set -u
python scripts/quick_bench.py --kernels 6 --reps 4 --lookups 8500000 XSB200_DENSE_MIN=64 XSB200_DENSE_MIN=48 XSB200_DENSE_MIN=32 XSB200_DENSE_MIN=24 XSB200_DENSE_MIN=16 XSB200_DENSE_MIN=32,XSB200_ARITH=fused 2>&1 | tail -6
python scripts/quick_bench.py --kernels 6 --reps 4 --lookups 4250000 XSB200_DENSE_MIN=64 XSB200_DENSE_MIN=32 XSB200_DENSE_MIN=16 XSB200_DENSE_MIN=8 2>&1 | tail -4
python scripts/e2e_bench.py "" XSB200_ARITH=fused XSB200_DENSE_MIN=48 XSB200_DENSE_MIN=32 XSB200_DENSE_MIN=24 XSB200_DENSE_MIN=32,XSB200_ARITH=fused XSB200_DENSE_MIN=32,XSB200_E2E_CHUNKS=3 XSB200_DENSE_MIN=24,XSB200_E2E_CHUNKS=3,XSB200_ARITH=fused XSB200_DENSE_MIN=16,XSB200_E2E_CHUNKS=4,XSB200_ARITH=fused 2>&1 | tail -10
