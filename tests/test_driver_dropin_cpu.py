"""The UNMODIFIED reference drivers on top of the drop-in overlay, on CPU: `train/train_own_forget_cl.py` (continual, engine_cl) and
`train/train_own_forget.py` (single step, engine) must run everything that is host-side
orchestration -- argument parsing, config, the synthetic ImageFolder split into forget / remain sets, the loaders, the eager construction of
every backbone (this repo's ViT_face and ModifiedViT included, with the torchvision ViT-B/16 checkpoint pre-seeded), `mark_only_lora_as_trainable`,
the timm-built AdamW -- and then stop exactly where the first engine call happens, loudly, because there is no CPU fallback
(landmines 1-7).  Needs the reference tree (authoring container); skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


CL_ARGS = ["-b", "4", "-w", "cpu", "-d", "casia100", "-n", "VIT", "-e", "2", "-head", "CosFace", "--warmup-epochs", "0", "--lr", "1e-2",
           "--num_workers", "0", "--lora_rank", "8", "--decay-epochs", "100", "--vit_depth", "6", "--num_of_first_cls", "80", "--per_forget_cls", "5",
           "--BND", "105", "--beta", "0.15", "--alpha", "0.0001", "--min-lr", "1e-5", "--num_tasks", "4", "--wandb_group", "t",
           "--cl_beta_list", "0.3", "0.4", "0.28", "0.2", "--wandb_offline"]                                  # scripts/run_cl_forget.sh:225-233
SINGLE_ARGS = ["-b", "4", "-w", "cpu", "-d", "casia100", "-n", "VIT", "-e", "2", "-head", "CosFace", "--grouping", "block", "--data_ratio", "0.5",
               "--alpha_epoch", "1", "--warmup-epochs", "0", "--lr", "1e-2", "--num_workers", "0", "--lora_rank", "8", "--decay-epochs", "2",
               "--wandb_group", "t", "--vit_depth", "6", "--num_of_first_cls", "80", "--per_forget_cls", "20", "--BND", "110", "--beta", "0.15",
               "--alpha", "0.01", "--min-lr", "1e-5", "--few_shot", "--few_shot_num", "2", "--wandb_offline"]  # scripts/run_forget.sh


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train", "train_own_forget_cl.py")), reason="reference tree not present")
@pytest.mark.parametrize("driver,args,module", [("train_own_forget_cl.py", CL_ARGS, "engine_cl.py"), ("train_own_forget.py", SINGLE_ARGS, "engine.py")])
def test_unmodified_driver_runs_up_to_the_first_engine_call(tmp_path, driver, args, module):
    import numpy as np
    import torch
    import torchvision
    from PIL import Image
    rng = np.random.default_rng(0)
    data = tmp_path / "data" / "faces_webface_112x112_sub100_train_test"            # config.py:32
    for split, n in (("train", 3), ("test", 2)):
        for c in range(100):
            d = data / split / f"{c:04d}"
            d.mkdir(parents=True)
            for i in range(n):
                Image.fromarray(rng.integers(0, 256, (112, 112, 3), dtype=np.uint8)).save(d / f"{i}.jpg")
    ckpt = tmp_path / "torch_home" / "hub" / "checkpoints"
    ckpt.mkdir(parents=True)
    # landmine 1: BACKBONE_DICT builds vit_b_16(weights=IMAGENET1K_V1) even for -n VIT; the file hash is only checked on download
    torch.save(torchvision.models.vit_b_16(weights=None).state_dict(), ckpt / "vit_b_16-c867db91.pth")
    env = dict(os.environ, TORCH_HOME=str(tmp_path / "torch_home"), WANDB_MODE="offline", WANDB_DIR=str(tmp_path),
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "gs-lora_b200"), os.path.join(ROOT, "oracle", "shims"), REF]))
    cmd = [sys.executable, "-u", os.path.join(REF, "train", driver)] + args + ["--outdir", str(tmp_path / "out")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=tmp_path, env=env)
    log = out.stdout + out.stderr
    assert out.returncode != 0
    assert "Use LoRA in Transformer FFN" in log, log[-3000:]
    if module == "engine_cl.py":
        assert "Optimizer Generated" in log
        for i in range(6):
            for t in ("net.0.lora_A", "net.0.lora_B", "net.3.lora_A", "net.3.lora_B"):
                assert f"transformer.layers.{i}.1.fn.fn.{t} True" in log           # the driver's "Learnable parameters" listing
    # the first forward of the run is engine_cl.eval_data of THIS repo, and without a GPU it refuses instead of falling back
    assert os.path.join("gs-lora_b200", module) in log and "in eval_data" in log
    assert "GslError" in log and "no CPU fallback" in log
