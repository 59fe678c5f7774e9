#!/bin/bash
# visit r02l: attention forward with P in tensor memory (TS-mode P V), ping-pong worker groups, double-buffered K / V
set -u
OUT=gpurun_out; TAG=r02l; mkdir -p $OUT
for shp in "2 197 8" "3 17 8" "2 50 4" "1 128 2" "5 129 8" "4 208 8" "64 197 8"; do timeout 120 python scripts/dev_attn_one.py $shp 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rfE -x -k "attention" 2>&1 | grep -v "^$" | tail -5
REPS=20 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print("split", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "other", d["other_precision_mode"])
PY
tail -3 $OUT/${TAG}_bench.err; grep -E "grads all|worst" $OUT/${TAG}_pytest_gpu.log | cut -c1-160 | head -30
