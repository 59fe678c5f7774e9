"""Golden for the epoch loop: the UNMODIFIED reference `engine_cl.train_one_epoch` (engine_cl.py:12-244) run on CPU over seeded list loaders with
the unmodified reference ViT_face and a timm-restated AdamW.  Liberties (host plumbing only, arithmetic untouched): `Tensor.cuda` -> identity,
`data_prefetcher` -> a CPU iterator with the same next() contract (util/data_prefetcher.py:44-58 needs a CUDA stream), wandb disabled.

    python tests/golden/make_golden_epoch.py        # authoring container only (needs /root/reference)
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
import engine_cl as ref_cl  # noqa: E402  (reference)
import util.utils as ref_utils  # noqa: E402  (reference)
from vit_pytorch_face import ViT_face  # noqa: E402  (reference)
import loralib as lora  # noqa: E402
from timm.optim import create_optimizer  # noqa: E402

from oracle.vit_oracle import TINY, VitConfig, init_state_dict, lora_param_list  # noqa: E402


class CpuPrefetcher:
    """same contract as util.data_prefetcher.data_prefetcher: next() -> (samples, targets) or (None, None) when exhausted"""

    def __init__(self, loader, device, prefetch=True):
        self.it = iter(loader)

    def next(self):
        return next(self.it, (None, None))


def loaders(cfg, seed, n_remain, n_forget):
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    remain = [(torch.rand(4, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (4,), generator=g)) for _ in range(n_remain)]
    forget = [(torch.rand(3, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (3,), generator=g)) for _ in range(n_forget)]
    return remain, forget


def meter_state(m):
    return dict(val=float(m.val), avg=float(m.avg), sum=float(m.sum), count=int(m.count))


def run(use_proto):
    cfg = VitConfig(**{**TINY.to_dict(), "depth": 6})
    seed = 51 + int(use_proto)
    sd = init_state_dict(cfg, seed=seed)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank)
    m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=1e-2, BND=105.0, w_pf=0.5 if use_proto else 0.25, w_pr=0.7 if use_proto else 0.0, BND_pro=18.0)
    opt = create_optimizer(types.SimpleNamespace(lr=hp["lr"], weight_decay=hp["wd"], opt_eps=1e-8, opt_betas=None, opt="adamw"), m)
    remain, forget = loaders(cfg, seed + 100, 7, 3)
    g = torch.Generator().manual_seed(seed + 200)
    protos = torch.randn(cfg.num_class, cfg.dim, generator=g)
    proto_dict = {i: protos[i] for i in range(cfg.num_class)}
    meters = [ref_utils.AverageMeter() for _ in range(8)]
    lf, lr_, lt, ls, tf, tr, lpf, lpr = meters
    run_cfg = {"DATA_ROOT": "./data/faces_webface_112x112_sub100_train_test/", "BND_pro": hp["BND_pro"], "WORK_PATH": "/tmp", "BACKBONE_NAME": "VIT",
               "MULTI_GPU": False}
    ret = ref_cl.train_one_epoch(m, forget, remain, torch.device("cpu"), torch.nn.CrossEntropyLoss(), opt, 0, lf, lr_, lt, ls, tf, tr,
                                 hp["beta"], hp["alpha"], hp["BND"], 0, None, None, 0.0, 0.0, run_cfg, 2, use_proto, proto_dict, hp["w_pf"], hp["w_pr"],
                                 lpf, lpr)
    names = ["losses_forget", "losses_remain", "top1_forget", "top1_remain", "losses_total", "losses_structure", "losses_prototype_forget",
             "losses_prototype_remain"]
    return dict(cfg=cfg.to_dict(), seed=seed, hp=hp, use_proto=use_proto, loader_seed=seed + 100, n_remain=7, n_forget=3, prototypes=protos,
                batch=int(ret[0]), highest_H_mean=float(ret[1]), meters={n: meter_state(x) for n, x in zip(names, ret[2:])},
                params_after={n: m.get_parameter(n).detach().clone() for n in lora_param_list(cfg)},
                state_dict_checksum={k: float(v.double().abs().sum()) for k, v in sd.items()})


def main():
    assert ref_cl.__file__.startswith(REF), ref_cl.__file__
    ref_cl.data_prefetcher = CpuPrefetcher
    wandb.init(mode="disabled")
    gold = {"plain": run(False), "proto": run(True)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny6_epoch.pt")
    torch.save(gold, path)
    for k, v in gold.items():
        print(k, v["batch"], {n: round(x["avg"], 5) for n, x in v["meters"].items()})
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
