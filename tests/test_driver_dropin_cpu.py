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


IMAGENET_ARGS = ["-b", "4", "-w", "cpu", "-d", "imagenet100", "-n", "VIT_B16", "-e", "2", "-head", "CosFace", "--warmup-epochs", "0", "--lr", "1e-2",
                 "--num_workers", "0", "--lora_rank", "8", "--decay-epochs", "100", "--vit_depth", "6", "--num_of_first_cls", "80",
                 "--per_forget_cls", "5", "--data_ratio", "0.5", "--BND", "105", "--beta", "0.15", "--alpha", "0.0001", "--min-lr", "1e-5",
                 "--num_tasks", "4", "--wandb_group", "t", "--cl_beta_list", "0.2", "0.25", "0.25", "0.25", "--wandb_offline"]   # scripts/run_cl_forget_image.sh:15-21


def _imagenet100_tree(tmp_path, rng):
    """data/imagenet100/{train,test}/<wnid>/, the 1000-name class list and the 'missing 900' validation split (train_own_forget_cl.py:135-178)"""
    import numpy as np
    from PIL import Image
    names = [f"n{10000000 + i:08d}" for i in range(1000)]
    root = tmp_path / "data" / "imagenet100"
    root.mkdir(parents=True)
    (root / "imagenet_folder_names.txt").write_text("\n".join(names) + "\n")
    for split, n in (("train", 2), ("test", 1)):
        for c in names[::10]:
            d = root / split / c
            d.mkdir(parents=True)
            for i in range(n):
                Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(d / f"{i}.jpg")
    for c in names[1:40:10]:
        d = tmp_path / "data" / "imagenet_val_split" / "nonexist" / c
        d.mkdir(parents=True)
        Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(d / "0.jpg")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train", "train_own_forget_cl.py")), reason="reference tree not present")
@pytest.mark.parametrize("driver,args,module", [("train_own_forget_cl.py", CL_ARGS, "engine_cl.py"), ("train_own_forget.py", SINGLE_ARGS, "engine.py"),
                                                ("train_own_forget_cl.py", IMAGENET_ARGS, "engine_cl.py")])
def test_unmodified_driver_runs_up_to_the_first_engine_call(tmp_path, driver, args, module):
    import numpy as np
    import torch
    import torchvision
    from PIL import Image
    rng = np.random.default_rng(0)
    imagenet = "imagenet100" in args
    if imagenet:
        _imagenet100_tree(tmp_path, rng)
    data = tmp_path / "data" / "faces_webface_112x112_sub100_train_test"            # config.py:32
    for split, n in (() if imagenet else (("train", 3), ("test", 2))):
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
    if imagenet:        # the pretrained ViT-B/16 checkpoint doubles as the -r resume file (it restores the MLP weights replace_ffn_with_lora re-initialised)
        cmd += ["-r", str(ckpt / "vit_b_16-c867db91.pth")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=tmp_path, env=env)
    log = out.stdout + out.stderr
    assert out.returncode != 0
    if imagenet:    # config 4: ModifiedViT + replace_ffn_with_lora are built, the checkpoint loads with only lora_* keys missing, first eval is ours
        assert "VIT_B16 Backbone Generated" in log and "Loading Backbone Checkpoint" in log and "Wrong resume" not in log, log[-3000:]
    else:
        assert "Use LoRA in Transformer FFN" in log, log[-3000:]
    if module == "engine_cl.py" and not imagenet:
        assert "Optimizer Generated" in log
        for i in range(6):
            for t in ("net.0.lora_A", "net.0.lora_B", "net.3.lora_A", "net.3.lora_B"):
                assert f"transformer.layers.{i}.1.fn.fn.{t} True" in log           # the driver's "Learnable parameters" listing
    # the first forward of the run is engine_cl.eval_data of THIS repo, and without a GPU it refuses instead of falling back
    assert os.path.join("gs-lora_b200", module) in log and "in eval_data" in log
    assert "GslError" in log and "no CPU fallback" in log
