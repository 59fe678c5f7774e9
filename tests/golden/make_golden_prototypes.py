"""Golden vector for the class-prototype path: the UNMODIFIED reference `util.utils.calculate_prototypes` (util/utils.py:502-549) run on the
unmodified reference `ViT_face` (eval mode, CPU, `Tensor.cuda` patched to the identity as in make_golden.py) over a seeded TensorDataset.

    python tests/golden/make_golden_prototypes.py        # authoring container only (needs /root/reference)
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF, ROOT]
os.environ.setdefault("WANDB_MODE", "disabled")

import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
_ii = types.ModuleType("image_iter")
_ii.CustomSubset = type("CustomSubset", (), {})
sys.modules["image_iter"] = _ii

import util.utils as ref_utils  # noqa: E402  (reference)
from vit_pytorch_face import ViT_face  # noqa: E402  (reference)
import loralib as lora  # noqa: E402

from oracle.vit_oracle import TINY, VitConfig, init_state_dict  # noqa: E402


def main():
    assert ref_utils.__file__.startswith(REF), ref_utils.__file__
    cfg = VitConfig(**{**TINY.to_dict(), "depth": 6})
    seed, N, bs = 21, 29, 8
    sd = init_state_dict(cfg, seed=seed)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank)
    m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    g = torch.Generator().manual_seed(seed + 1)
    imgs = torch.rand(N, 3, cfg.image_size, cfg.image_size, generator=g)
    labs = torch.randint(0, 5, (N,), generator=g)          # 5 of the 10 classes occur
    protos = ref_utils.calculate_prototypes(m, torch.utils.data.TensorDataset(imgs, labs), batch_size=bs, device="cpu")
    gold = dict(cfg=cfg.to_dict(), seed=seed, batch_size=bs, images=imgs, labels=labs,
                prototypes={int(k): v.clone() for k, v in protos.items()},
                state_dict_checksum={k: float(v.double().abs().sum()) for k, v in sd.items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny6_prototypes.pt")
    torch.save(gold, path)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB", sorted(gold["prototypes"]))


if __name__ == "__main__":
    main()
