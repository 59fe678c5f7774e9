#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02h; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error|2 ranks" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 300 python scripts/dev_gaps.py > $OUT/${TAG}_gaps.log 2>&1; cat $OUT/${TAG}_gaps.log | tail -15
