#!/bin/bash
# visit r02r: precision mode split8 end to end -- full GPU suite with it as the process default, bench with all three modes
set -u
OUT=gpurun_out; TAG=r02r; mkdir -p $OUT
GSLORA_PRECISION=split8 timeout 1200 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu_split8.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu_split8.log | cut -c1-300 | tail -12
grep -E "grads all|worst|window|norms" $OUT/${TAG}_pytest_gpu_split8.log | cut -c1-200 | head -50
timeout 600 python bench.py --precision split8 --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
print(d["config"]["precision"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "others", d["other_precision_modes"])
print("roof", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["ms_per_launch_pair"], [ (o["kernel"][-40:], o["ms_per_launch_pair"]) for o in d["roofline"]["other_modes"]])
PY
tail -3 $OUT/${TAG}_bench.err
