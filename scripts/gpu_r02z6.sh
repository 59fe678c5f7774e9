#!/bin/bash
set -u
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-u8-leg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['precision'], d['value'], d['ms_per_step']); print(d['roofline']['kernel']); print([o['kernel'][:60] for o in d['roofline']['other_modes']])"
