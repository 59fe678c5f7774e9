#!/bin/bash
# visit r02y (final build of round 2): full suite, smoke, the driver's bench commands (repo arm + reference arm), launch list, ncu full, sanitizers
set -u
OUT=gpurun_out; TAG=r02y; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print(d["config"]["precision"], d["value"], d["ms_per_step"], "steps", d["steps"], "e2e", d["e2e"]["value"], "u8", d["e2e"]["uint8_pipeline"]["value"], "launches", d["gpu_launches"], "others", d["other_precision_modes"])
print("roof", d["roofline"]["frac"], d["roofline"]["executed_frac"], d["roofline"]["ms_per_launch_pair"], "gpu_ref", d["gpu_reference"]["ms_per_step"], d["gpu_reference"]["ratio"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
tail -3 $OUT/${TAG}_bench.err
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 400 $OUT/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference > $OUT/${TAG}_launches.log 2>&1; tail -1 $OUT/${TAG}_launches.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attention_|lora_side' -c 22 -f \
  -o $OUT/${TAG}_full python scripts/dev_prof.py split8 attn skinny > $OUT/${TAG}_full.log 2>&1; tail -2 $OUT/${TAG}_full.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_sanitizer_memcheck.log; tail -3 $OUT/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_sanitizer_racecheck.log; tail -3 $OUT/${TAG}_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> $OUT/${TAG}_sanitizer_synccheck.log; tail -3 $OUT/${TAG}_sanitizer_synccheck.log
timeout 300 python scripts/dev_gaps.py 2>&1 | grep -E "wall|sync" | tee $OUT/${TAG}_gaps.log
