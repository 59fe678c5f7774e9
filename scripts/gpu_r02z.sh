#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02z; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "split8" 2>&1 | tail -2
for pdl in 0 1; do GSLORA_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg --single-mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"; done
