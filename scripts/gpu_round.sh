#!/bin/bash
# One GPU-box visit: parity tests, bench line, launch list (ncu, one metric) and a full ncu capture of the hot kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [tests|bench|launches|full ...]
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)
      GSLORA_SPEEDUP_OUT=$OUT/${TAG}_speedup.json timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
      tail -5 $OUT/${TAG}_pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
      timeout 300 python scripts/dev_gaps.py > $OUT/${TAG}_gaps.log 2>&1; cat $OUT/${TAG}_gaps.log ;;
    kernels)
      REPS=10 timeout 300 python scripts/dev_prof.py > $OUT/${TAG}_kernels.log 2>&1; cat $OUT/${TAG}_kernels.log ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-u8-leg > $OUT/${TAG}_launches.log 2>&1; tail -2 $OUT/${TAG}_launches.log ;;
    full)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attention_|lora_side' -c 16 -f \
        -o $OUT/${TAG}_full python scripts/dev_prof.py gelu res f16 attn skinny > $OUT/${TAG}_full.log 2>&1; tail -3 $OUT/${TAG}_full.log ;;
  esac
done
ls -la $OUT
