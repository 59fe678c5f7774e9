#!/bin/bash
# visit r02k: quick wins (parallel split reduction, strip-staged patchify, multi-image head kernels) + attention stage knock-outs
set -u
OUT=gpurun_out; TAG=r02k; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -rfE -x 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
for d in 0 1 2 4 6 8 16 20; do echo "== GSL_ATTN_DBG=$d"; GSL_ATTN_DBG=$d REPS=10 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2; done > $OUT/${TAG}_attn_dbg.log 2>&1; cat $OUT/${TAG}_attn_dbg.log
REPS=10 timeout 120 python scripts/dev_prof.py skinny > $OUT/${TAG}_skinny.log 2>&1; cat $OUT/${TAG}_skinny.log
timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print("split", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "other", d["other_precision_mode"])
PY
tail -3 $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference > $OUT/${TAG}_launches.log 2>&1; tail -2 $OUT/${TAG}_launches.log
