#!/bin/bash
# visit r02p: attention forward (TMEM-resident P), delta from the dO GEMM epilogue, attention backward reading it
set -u
OUT=gpurun_out; TAG=r02p; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rfE -x -k "attention or rowdot or gemm_epilogues" 2>&1 | grep -v "^$" | tail -5
REPS=20 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
grep -E "grads all|worst" $OUT/${TAG}_pytest_gpu.log | cut -c1-160 | head -30
timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print("split", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "other", d["other_precision_mode"])
PY
tail -3 $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference > $OUT/${TAG}_launches.log 2>&1; tail -1 $OUT/${TAG}_launches.log | cut -c1-200
