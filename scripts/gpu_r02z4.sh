#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02z4; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cls_attention' -c 6 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference 2>/dev/null | grep -E "cls_attention" | cut -d'"' -f10,30- | head -6
timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg --single-mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['precision'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
