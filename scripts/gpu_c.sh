#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "energy_bands or all_variants" 2>&1 | tail -8
