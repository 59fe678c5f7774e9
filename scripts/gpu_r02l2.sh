#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02l2; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
grep -E "grads all|worst" $OUT/${TAG}_pytest_gpu.log | cut -c1-200 | head -60
