#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -k "split8" 2>&1 | grep -E "split8|passed|failed|Error" | cut -c1-200
