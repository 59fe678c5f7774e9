#!/bin/bash
# round 2, final build at 8 GPUs of one box: the headline workload and the two 224x224 configurations (precision split8)
set -u
OUT=gpurun_out; TAG=r02x8; mkdir -p $OUT
COMMON="--steps 15 --warmup 3 --no-cpu-baseline --no-u8-leg --single-mode --no-gpu-reference"
for wl in p8s8_bs512 vitb16_bs48 vitl16_bs32; do
  f=$OUT/${TAG}_${wl}_n8
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --workload $wl $COMMON > $f.json 2> $f.err
  python - "$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["n_gpus"], "gpus", d["value"], "img/s", d["ms_per_step"], "ms/step", "e2e", d["e2e"]["value"], d["config"]["precision"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
