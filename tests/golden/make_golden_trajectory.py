"""Golden for FREE-RUNNING trajectories: the UNMODIFIED reference `engine.train_one_epoch` (engine.py:13-434) on CPU for 50 consecutive
optimizer steps with nothing teacher-forced -- loralib's own start (lora_B = 0), lr 1e-2, the ALPHA_EPOCH switch (engine.py:82-90: epoch 0 =
15 steps without the structure term, epoch 1 = 35 steps with alpha = 2, large enough that the group lasso collapses most blocks' LoRA groups
while the data term keeps others alive) and a forget bound the run actually reaches (BND = 36: relu(BND - CE_f) switches on and off).
Recorded: every wandb.log record of the loop (the running meters every 5 steps), the LoRA parameters after each epoch and the per-block
group norms.  Same liberties as make_golden_epoch.py (host plumbing only).

    python tests/golden/make_golden_trajectory.py        # authoring container only (needs /root/reference)
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

from oracle.vit_oracle import TINY, VitConfig, init_state_dict, lora_names, lora_param_list  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_epoch import CpuPrefetcher  # noqa: E402

HP = dict(lr=1e-2, wd=0.05, beta=1.0, alpha=2.0, BND=36.0, alpha_epoch=1, steps=(15, 35), batch=8, seed=77)


def trajectory_loaders(cfg, seed, n, bs, distinct=2):
    """n (remain, forget) batch pairs of bs images, cycling through `distinct` different pairs (a small set the run can actually fit, so the
    data term has a direction that competes with the group lasso) -- shared with tests/test_trajectory*.py"""
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    rb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    fb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    return [rb[i % distinct] for i in range(n)], [fb[i % distinct] for i in range(n)]


def group_norms(params, cfg):
    return [float(torch.sqrt(sum((params[n].double() ** 2).sum() for n in grp))) for grp in lora_names(cfg)]


def main():
    assert ref_engine.__file__.startswith(REF), ref_engine.__file__
    ref_engine.data_prefetcher = CpuPrefetcher
    wandb.init(mode="disabled")
    records = []
    ref_engine.wandb = types.SimpleNamespace(log=lambda d: records.append({k: float(v) for k, v in d.items()}))
    cfg = VitConfig(**{**TINY.to_dict(), "depth": 6})
    sd = init_state_dict(cfg, seed=HP["seed"], lora_b_std=0.0)            # loralib.Linear.reset_parameters: lora_B = 0
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank)
    m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    opt = create_optimizer(types.SimpleNamespace(lr=HP["lr"], weight_decay=HP["wd"], opt_eps=1e-8, opt_betas=None, opt="adamw"), m)
    run_cfg = {"few_shot": False, "ALPHA_EPOCH": HP["alpha_epoch"], "NUM_LAYERS": cfg.depth, "GROUP_TYPE": "block", "GROUP_POS": "FFN",
               "WORK_PATH": "/tmp", "BACKBONE_NAME": "VIT", "MULTI_GPU": False}
    names = lora_param_list(cfg)
    batch, after, norms = 0, [], []
    for epoch, n in enumerate(HP["steps"]):
        remain, forget = trajectory_loaders(cfg, HP["seed"] + 10, n, HP["batch"])        # the same small set in both epochs
        meters = [ref_utils.AverageMeter() for _ in range(8)]
        ret = ref_engine.train_one_epoch(m, forget, remain, torch.device("cpu"), torch.nn.CrossEntropyLoss(), opt, epoch, meters[0], meters[1],
                                         meters[2], meters[3], meters[4], meters[5], HP["beta"], HP["alpha"], HP["BND"], batch, None, None, 0.0, 0.0,
                                         run_cfg, losses_prototype_forget=meters[6], losses_prototype_remain=meters[7])
        batch = int(ret[0])
        params = {n_: m.get_parameter(n_).detach().clone() for n_ in names}
        after.append(params)
        norms.append(group_norms(params, cfg))
    gold = dict(cfg=cfg.to_dict(), hp=HP, records=records, params_after=after, group_norms=norms, batch=batch,
                state_dict_checksum={k: float(v.double().abs().sum()) for k, v in sd.items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny6_trajectory.pt")
    torch.save(gold, path)
    for r in records:
        print({k.replace("epoch_", ""): round(v, 4) for k, v in r.items() if "prototype" not in k})
    for e, nn_ in enumerate(norms):
        print("group norms after epoch", e, [round(x, 4) for x in nn_])
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
