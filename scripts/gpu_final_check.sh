#!/bin/bash
set -u
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-260
