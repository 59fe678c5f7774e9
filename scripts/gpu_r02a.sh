#!/bin/bash
# round 2, visit a: split-precision parity + timing
set -u
OUT=gpurun_out; TAG=r02a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 600 python tests/dev_parity.py > $OUT/${TAG}_parity.log 2>&1; tail -12 $OUT/${TAG}_parity.log
timeout 1500 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; tail -40 $OUT/${TAG}_pytest_gpu.log
REPS=10 timeout 300 python scripts/dev_prof.py gelu res f16 split > $OUT/${TAG}_kernels.log 2>&1; cat $OUT/${TAG}_kernels.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
