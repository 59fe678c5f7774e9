#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02c; mkdir -p $OUT
timeout 300 python scripts/dev_op_errors.py > $OUT/${TAG}_op_errors.log 2>&1; cat $OUT/${TAG}_op_errors.log
timeout 900 python -m pytest tests/test_trajectory_gpu.py tests/test_driver_replay_gpu.py tests/test_kernels_gpu.py -m gpu -q -rfE -s -x 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest.log; tail -60 $OUT/${TAG}_pytest.log
