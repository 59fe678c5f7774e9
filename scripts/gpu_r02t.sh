#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02w; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -2
REPS=20 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2
REPS=10 timeout 300 python scripts/dev_prof.py split8 > $OUT/${TAG}_kernels.log 2>&1; cat $OUT/${TAG}_kernels.log
timeout 900 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print(d["config"]["precision"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "others", d["other_precision_modes"])
print("roof", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["ms_per_launch_pair"])
PY
tail -3 $OUT/${TAG}_bench.err
