#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02d; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "trajectory|P8S8 bs|tv family|passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-400 | tail -60
