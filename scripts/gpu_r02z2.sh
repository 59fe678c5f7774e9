#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02z2; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-u8-leg --single-mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['precision'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
