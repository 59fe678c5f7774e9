"""The UNMODIFIED continual driver, START TO FINISH, in the authoring container: `train/train_own_forget_cl.py` (2 tasks, prototypes, EMA -- the
flags of scripts/run_cl_forget.sh:225-233) on top of the drop-in overlay, with the native engine replaced INSIDE THE SUBPROCESS by an
oracle-backed CPU stand-in (tests/cpu_engine/sitecustomize.py: test infrastructure, product code untouched -- without it the run stops at the
first engine call, tests/test_driver_dropin_cpu.py).  Everything after that first call is therefore exercised against the real driver: eval x4
(+ old), train_one_epoch, the EMA deep copies and their evaluation, the norm report, eval-mode task checkpoints, reload +
reinitialize_lora_parameters + a new optimizer (fused moments reset) for task 1, the final old-class evaluation.

The recorded sequence of drop-in entry points is then compared with the one tests/driver_replay.py produces under the same stand-in: the
replay harness -- which is what runs this sequence on the GPU box, where the reference tree does not exist (tests/test_driver_replay_gpu.py)
-- re-enacts the real driver call for call.  Needs the reference tree; skipped elsewhere."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

ARGS = ["-b", "8", "-w", "cpu", "-d", "casia100", "-n", "VIT", "-e", "1", "-head", "CosFace", "--warmup-epochs", "0", "--lr", "1e-2", "--num_workers", "0",
        "--lora_rank", "8", "--decay-epochs", "100", "--vit_depth", "2", "--num_of_first_cls", "90", "--per_forget_cls", "5", "--BND", "105", "--beta", "0.15",
        "--alpha", "0.0001", "--min-lr", "1e-5", "--num_tasks", "2", "--wandb_group", "t", "--cl_beta_list", "0.3", "0.4", "--wandb_offline", "--prototype",
        "--pro_f_weight", "0.017", "--pro_r_weight", "0.01", "--average_weight", "--ema_epoch", "0", "--ema_decay", "0.9", "--cl_prof_list", "0.015", "0.06",
        "--BND_pro", "50"]

EXPECTED = (["calculate_prototypes"] + ["eval_data"] * 4 + ["train_one_epoch", "reset_optimizer"] + ["eval_data"] * 2 + ["get_norm_of_lora", "save_task_checkpoint"] +
            ["reinitialize_lora_parameters", "calculate_prototypes"] + ["eval_data"] * 5 + ["train_one_epoch", "reset_optimizer"] + ["eval_data"] * 2 +
            ["get_norm_of_lora", "save_task_checkpoint", "eval_data"])


def _env(tmp_path, trace):
    return dict(os.environ, GSLORA_CPU_ORACLE_ENGINE="1", GSLORA_CUDA_GRAPH="0", GSLORA_TRACE=str(trace), TORCH_HOME=str(tmp_path / "torch_home"), WANDB_MODE="offline",
                WANDB_DIR=str(tmp_path),
                PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "cpu_engine"), os.path.join(ROOT, "gs-lora_b200"), os.path.join(ROOT, "oracle", "shims"), REF]))


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train", "train_own_forget_cl.py")), reason="reference tree not present")
def test_unmodified_driver_runs_two_tasks_to_completion_and_the_replay_matches_its_call_sequence(tmp_path):
    import numpy as np
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
    torch.save(torchvision.models.vit_b_16(weights=None).state_dict(), ckpt / "vit_b_16-c867db91.pth")      # landmine 1
    trace = tmp_path / "trace_driver.txt"
    out = subprocess.run([sys.executable, "-u", os.path.join(REF, "train", "train_own_forget_cl.py")] + ARGS + ["--outdir", str(tmp_path / "out")],
                         capture_output=True, text=True, timeout=1500, cwd=tmp_path, env=_env(tmp_path, trace))
    log = out.stdout + out.stderr
    assert out.returncode == 0, log[-4000:]
    assert "task:1" in log and "start one stage forget remain training" in log and "Test old-1 Accuracy" in log
    calls = trace.read_text().split()
    assert calls == EXPECTED, calls
    # task checkpoints: the reference's key set (weights + lora_A / lora_B), written in eval mode => merged FFN weights
    run_dirs = [os.path.join(dp, "task-level") for dp, dn, _ in os.walk(tmp_path / "out") if "task-level" in dn]
    assert len(run_dirs) == 1
    sd0, sd1 = [torch.load(os.path.join(run_dirs[0], f"Backbone_task_{t}.pth")) for t in (0, 1)]
    keys = set(sd0)
    assert keys == set(sd1) and sum("lora_A" in k for k in keys) == 4 and "loss.weight" in keys and "transformer.layers.1.1.fn.fn.net.3.lora_B" in keys
    assert float(sd0["transformer.layers.0.1.fn.fn.net.0.lora_B"].abs().max()) > 0          # task 0 trained (lora_B starts at zero)
    # task 1 restarted its LoRA from the merged task-0 model: lora_A was re-drawn, and the frozen FFN weight of task 1's checkpoint is task 0's merged
    # weight plus task 1's own delta
    A0, A1 = sd0["transformer.layers.0.1.fn.fn.net.0.lora_A"], sd1["transformer.layers.0.1.fn.fn.net.0.lora_A"]
    assert not torch.allclose(A0, A1)
    w0, w1 = sd0["transformer.layers.0.1.fn.fn.net.0.weight"], sd1["transformer.layers.0.1.fn.fn.net.0.weight"]
    delta1 = (sd1["transformer.layers.0.1.fn.fn.net.0.lora_B"] @ A1) / 8.0
    assert float((w1 - w0 - delta1).norm() / delta1.norm()) < 1e-3
    # the replay harness, same stand-in, same settings: identical call sequence
    trace2 = tmp_path / "trace_replay.txt"
    rep = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "tests", "cpu_engine", "run_replay.py"), str(tmp_path / "replay_out")],
                         capture_output=True, text=True, timeout=900, cwd=tmp_path, env=_env(tmp_path, trace2))
    assert rep.returncode == 0, (rep.stdout + rep.stderr)[-4000:]
    assert trace2.read_text().split() == calls
