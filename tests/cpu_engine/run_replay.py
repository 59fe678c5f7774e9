"""Runs tests/driver_replay.py on CPU under the oracle-engine stand-in (tests/cpu_engine/sitecustomize.py) with the settings of the driver run in
tests/test_driver_full_cpu.py, so that the two call traces (GSLORA_TRACE) can be compared."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from driver_replay import replay  # noqa: E402
import loralib as lora  # noqa: E402
from vit_pytorch_face import ViT_face  # noqa: E402

torch.manual_seed(0)
m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=100, image_size=112, patch_size=8, dim=512, depth=2, heads=8, mlp_dim=2048, dropout=0.1,
             emb_dropout=0.1, lora_rank=8)
lora.mark_only_lora_as_trainable(m)
out = replay(m, image_size=112, num_class=100, device=torch.device("cpu"), work_path=sys.argv[1], num_tasks=2, epochs=1, batch_size=8, first_cls=90,
             per_forget=5, per_class=3, prototype=True, average_weight=True, ema_epoch=0)
print("replay done", [t["steps"] for t in out["tasks"]])
