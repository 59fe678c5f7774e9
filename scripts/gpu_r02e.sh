#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02e; mkdir -p $OUT
timeout 900 python -m pytest tests/test_next_rows_gpu.py tests/test_driver_replay_gpu.py -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest.log; grep -E "attention LoRA|passed|failed|FAILED|Error|assert" $OUT/${TAG}_pytest.log | cut -c1-300 | tail -30
# hygiene: compute-sanitizer on the tiny step (smoke = forward, selective backward, fused optimizer of a depth-2 512-wide model)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_sanitizer_memcheck.log; tail -6 $OUT/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_sanitizer_racecheck.log; tail -6 $OUT/${TAG}_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> $OUT/${TAG}_sanitizer_synccheck.log; tail -4 $OUT/${TAG}_sanitizer_synccheck.log
# launch list (split mode, the default) and a full capture of the top kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference > $OUT/${TAG}_launches.log 2>&1; tail -2 $OUT/${TAG}_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attention_|lora_side|layernorm' -c 24 -f \
  -o $OUT/${TAG}_full python scripts/dev_prof.py split attn skinny > $OUT/${TAG}_full.log 2>&1; tail -3 $OUT/${TAG}_full.log
ls -la $OUT | tail -12
