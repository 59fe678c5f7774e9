"""Golden for the single-step twin: the UNMODIFIED reference `engine.train_one_epoch` (engine.py:13-434) on CPU -- the few-shot branch
(forget loader longer, cfg["few_shot"]: the forget loader drives, engine.py:53-57) with GROUP_TYPE "lora", and the ordinary branch with the
structure term gated off by ALPHA_EPOCH (engine.py:82-90) and GROUP_TYPE "matrix".  Same liberties as make_golden_epoch.py.

    python tests/golden/make_golden_epoch_single.py        # authoring container only (needs /root/reference)
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF, ROOT]
os.environ["WANDB_MODE"] = "disabled"

import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
_ii = types.ModuleType("image_iter")
_ii.CustomSubset = type("CustomSubset", (), {})
sys.modules["image_iter"] = _ii

import wandb  # noqa: E402
import engine as ref_engine  # noqa: E402  (reference)
import util.utils as ref_utils  # noqa: E402  (reference)
from vit_pytorch_face import ViT_face  # noqa: E402  (reference)
import loralib as lora  # noqa: E402
from timm.optim import create_optimizer  # noqa: E402

from oracle.vit_oracle import TINY, VitConfig, init_state_dict, lora_param_list  # noqa: E402
from make_golden_epoch import CpuPrefetcher, loaders, meter_state  # noqa: E402


def run(name, few_shot, n_remain, n_forget, alpha_epoch, group_type, epoch):
    cfg = VitConfig(**{**TINY.to_dict(), "depth": 3})
    seed = 61 + len(name)
    sd = init_state_dict(cfg, seed=seed)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank)
    m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=2e-2, BND=105.0)
    opt = create_optimizer(types.SimpleNamespace(lr=hp["lr"], weight_decay=hp["wd"], opt_eps=1e-8, opt_betas=None, opt="adamw"), m)
    remain, forget = loaders(cfg, seed + 100, n_remain, n_forget)
    meters = [ref_utils.AverageMeter() for _ in range(8)]
    run_cfg = {"few_shot": few_shot, "ALPHA_EPOCH": alpha_epoch, "NUM_LAYERS": cfg.depth, "GROUP_TYPE": group_type, "GROUP_POS": "FFN",
               "WORK_PATH": "/tmp", "BACKBONE_NAME": "VIT", "MULTI_GPU": False}
    ret = ref_engine.train_one_epoch(m, forget, remain, torch.device("cpu"), torch.nn.CrossEntropyLoss(), opt, epoch, meters[0], meters[1], meters[2],
                                     meters[3], meters[4], meters[5], hp["beta"], hp["alpha"], hp["BND"], 0, None, None, 0.0, 0.0, run_cfg,
                                     losses_prototype_forget=meters[6], losses_prototype_remain=meters[7])
    names = ["losses_forget", "losses_remain", "top1_forget", "top1_remain", "losses_total", "losses_structure", "losses_prototype_forget",
             "losses_prototype_remain"]
    return dict(cfg=cfg.to_dict(), seed=seed, hp=hp, loader_seed=seed + 100, n_remain=n_remain, n_forget=n_forget, run_cfg=run_cfg, epoch=epoch,
                batch=int(ret[0]), highest_H_mean=float(ret[1]), meters={n: meter_state(x) for n, x in zip(names, ret[2:])},
                params_after={n: m.get_parameter(n).detach().clone() for n in lora_param_list(cfg)},
                state_dict_checksum={k: float(v.double().abs().sum()) for k, v in sd.items()})


def main():
    assert ref_engine.__file__.startswith(REF), ref_engine.__file__
    ref_engine.data_prefetcher = CpuPrefetcher
    wandb.init(mode="disabled")
    gold = {"few_shot_lora": run("few_shot_lora", True, 2, 6, 0, "lora", 0),
            "gated_matrix": run("gated_matrix", False, 6, 2, 1, "matrix", 0),
            "open_block": run("open_block", False, 3, 4, 1, "block", 1)}       # forget loader longer but few_shot off; epoch == ALPHA_EPOCH: term on
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny3_epoch_single.pt")
    torch.save(gold, path)
    for k, v in gold.items():
        print(k, v["batch"], {n: round(x["avg"], 5) for n, x in v["meters"].items()})
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
