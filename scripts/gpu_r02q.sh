#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02q2; mkdir -p $OUT
REPS=10 timeout 300 python scripts/dev_prof.py split8 > $OUT/${TAG}_kernels.log 2>&1; cat $OUT/${TAG}_kernels.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "split8" 2>&1 | tail -2
