#!/bin/bash
# round 2, visit b: what sets the remaining per-tensor gradient error in split mode?
set -u
OUT=gpurun_out; TAG=r02b; mkdir -p $OUT
run() { echo "== $*"; env "$@" timeout 600 python tests/dev_parity.py 2>&1 | grep -v Warning | grep -v "return float"; }
{
run MODES=split SEEDS="1 2" VERBOSE=1
run MODES=split SEEDS="1 2" GSLORA_GRAD_SCALE=16384
run MODES=split SEEDS="1 2" GSLORA_GRAD_SCALE=65536
run MODES=split SEEDS="1 2" GSLORA_DXN32=1
run MODES=split SEEDS="1 2" GSLORA_DXN32=1 GSLORA_GRAD_SCALE=16384
run MODES=split SEEDS="1 2" B=128
run MODES=split SEEDS="1 2" B=128 GSLORA_GRAD_SCALE=65536
run MODES=split SEEDS="1 2" B=8
} > $OUT/${TAG}_parity.log 2>&1
cat $OUT/${TAG}_parity.log
