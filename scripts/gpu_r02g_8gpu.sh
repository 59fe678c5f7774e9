#!/bin/bash
# round 2, 8-GPU visit: 224x224 configurations (BASELINE configs 4 / 5) at 1 / 2 / 4 / 8 GPUs, alpha sweep at 8, P8S8 at 8, 2-rank driver replay
set -u
OUT=gpurun_out; TAG=r02g; mkdir -p $OUT
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
    print(sys.argv[1].split('/')[-1], d["n_gpus"], "gpus", d["value"], "img/s", d["ms_per_step"], "ms/step", "e2e", d["e2e"]["value"], "alpha", d["config"].get("alpha"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
nvidia-smi --query-gpu=index,name --format=csv | head -10 > $OUT/${TAG}_smi.txt
for n in 1 2 4 8; do run vitl16_bs32 $n; done
for n in 1 2 4 8; do run vitb16_bs48 $n; done
for a in 0 1e-3 1e-2; do run vitl16_bs32 8 --alpha $a; done
run p8s8_bs512 8
GSLORA_REPLAY_REF=$OUT/r02f_replay_ref.pt timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610 scripts/ddp_replay.py > $OUT/${TAG}_ddp_replay.log 2>&1; grep ddp_replay $OUT/${TAG}_ddp_replay.log | tail -2; tail -3 $OUT/${TAG}_ddp_replay.log | cut -c1-300
