#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02m; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_' -c 4 -f -o $OUT/${TAG}_attn python scripts/dev_prof.py attn > $OUT/${TAG}_attn.log 2>&1; tail -3 $OUT/${TAG}_attn.log
ls -la $OUT/${TAG}_attn.ncu-rep
