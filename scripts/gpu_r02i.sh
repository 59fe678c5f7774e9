#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02i; mkdir -p $OUT
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -rfE -s -k "graph or fused_step or pipelined or dropout" 2>&1 | grep -v "^$" | tail -30 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -rfE 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print("split", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "u8", d["e2e"]["uint8_pipeline"]["value"], "launches", d["gpu_launches"], "other", d["other_precision_mode"], "gpu_ref", d["gpu_reference"])
PY
tail -3 $OUT/${TAG}_bench.err
GSLORA_CUDA_GRAPH=0 timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-gpu-reference --no-u8-leg > $OUT/${TAG}_bench_nograph.json 2> $OUT/${TAG}_bench_nograph.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_nograph.json').read().strip().splitlines()[-1])
print("NO GRAPH split", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "other", d["other_precision_mode"])
PY
timeout 300 python scripts/dev_gaps.py 2>&1 | grep -E "wall|sync" 
