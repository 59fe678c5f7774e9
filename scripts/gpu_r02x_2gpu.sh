#!/bin/bash
# round 2, 2-GPU visit with the final default (precision split8): the driver's scaling command at N = 1, 2; 224x224 workloads at N = 1, 2; 2-rank driver replay
set -u
OUT=gpurun_out; TAG=r02x; mkdir -p $OUT
COMMON="--steps 15 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference"
run() {  # workload N extra...
  local wl=$1 n=$2; shift 2
  local f=$OUT/${TAG}_${wl}_n${n}$(echo "$*" | tr -d ' -')
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 --workload $wl $COMMON "$@" > $f.json 2> $f.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --workload $wl $COMMON "$@" > $f.json 2> $f.err; fi
  python - "$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["n_gpus"], "gpus", d["value"], "img/s", d["ms_per_step"], "ms/step", "e2e", d["e2e"]["value"], d["config"]["precision"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for n in 1 2; do run p8s8_bs512 $n; done
for n in 1 2; do run vitb16_bs48 $n; done
for n in 1 2; do run vitl16_bs32 $n; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610 scripts/ddp_replay.py > $OUT/${TAG}_ddp_replay.log 2>&1; grep ddp_replay $OUT/${TAG}_ddp_replay.log | tail -2; tail -2 $OUT/${TAG}_ddp_replay.log | cut -c1-300
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q 2>&1 | tail -2
# the driver's exact command shape at N = 2 (full default bench)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/${TAG}_driver_n2.json 2> $OUT/${TAG}_driver_n2.err; tail -c 600 $OUT/${TAG}_driver_n2.json; tail -2 $OUT/${TAG}_driver_n2.err
