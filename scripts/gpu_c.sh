#!/bin/bash
set -u
timeout 1200 python -m pytest tests -x -q -m "gpu and slow" -k billion 2>&1 | tail -4
