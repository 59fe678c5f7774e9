#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02n; mkdir -p $OUT
for shp in "2 197 8" "3 17 8" "2 50 4" "1 128 2" "5 129 8" "4 208 8" "64 197 8"; do timeout 120 python scripts/dev_attn_one.py $shp 2>&1 | tail -1; done
REPS=20 timeout 120 python scripts/dev_prof.py attn 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_fwd' -c 2 -f -o $OUT/${TAG}_attn python scripts/dev_prof.py attn > $OUT/${TAG}_attn.log 2>&1; tail -1 $OUT/${TAG}_attn.log
