#!/bin/bash
set -u
timeout 1500 python -m pytest tests -x -q -m "gpu and not slow" 2>&1 | tail -5
