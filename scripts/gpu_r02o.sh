#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02o; mkdir -p $OUT
for t in 1 0; do echo "== turns=$t"; GSL_ATTN_TURNS=$t REPS=20 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2 | head -1; GSL_ATTN_TURNS=$t timeout 120 python scripts/dev_attn_trace.py 2>&1 | tail -26; done > $OUT/${TAG}_trace.log 2>&1
cat $OUT/${TAG}_trace.log
