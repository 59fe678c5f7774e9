#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02f; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfE -s 2>&1 | grep -v "^$" > $OUT/${TAG}_pytest_gpu.log; grep -E "passed|failed|FAILED|Error" $OUT/${TAG}_pytest_gpu.log | cut -c1-300 | tail -12
timeout 600 python bench.py --workload vitb16_bs48 --steps 20 --warmup 3 > $OUT/${TAG}_bench_vitb16.json 2> $OUT/${TAG}_bench_vitb16.err; cut -c1-600 $OUT/${TAG}_bench_vitb16.json; tail -2 $OUT/${TAG}_bench_vitb16.err
timeout 600 python bench.py --workload vitl16_bs32 --steps 20 --warmup 3 > $OUT/${TAG}_bench_vitl16.json 2> $OUT/${TAG}_bench_vitl16.err; cut -c1-600 $OUT/${TAG}_bench_vitl16.json; tail -2 $OUT/${TAG}_bench_vitl16.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; cat $OUT/${TAG}_bench_reference.json; tail -4 $OUT/${TAG}_bench_reference.err
free -g | head -2; nproc
GSLORA_REPLAY_REF=$OUT/${TAG}_replay_ref.pt timeout 300 python scripts/ddp_replay.py 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
