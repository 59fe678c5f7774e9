#!/bin/bash
# last visit of round 2: the driver's bench command and the launch list on the final build (after the cls-attention change)
set -u
OUT=gpurun_out; TAG=r02z5; mkdir -p $OUT
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z5_bench.json').read().strip().splitlines()[-1])
print(d["config"]["precision"], d["value"], d["ms_per_step"], "steps", d["steps"], "e2e", d["e2e"]["value"], "u8", d["e2e"]["uint8_pipeline"]["value"], "launches", d["gpu_launches"], "others", d["other_precision_modes"])
print("roof", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["ms_per_launch_pair"], "gpu_ref", d["gpu_reference"]["ms_per_step"], d["gpu_reference"]["ratio"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference > $OUT/${TAG}_launches.log 2>&1; tail -1 $OUT/${TAG}_launches.log | cut -c1-100
