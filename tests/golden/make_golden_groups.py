"""Golden values for the group-lasso groupings: the UNMODIFIED reference `engine.get_structure_loss(model, num_layers, group_type, group_pos="FFN")`
(engine.py:532-687) and `util.cal_norm.get_norm_of_lora(..., group_type=...)` (util/cal_norm.py:4-146) on the unmodified reference ViT_face.

    python tests/golden/make_golden_groups.py        # authoring container only (needs /root/reference)
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

import engine as ref_engine  # noqa: E402  (reference)
import util.cal_norm as cal_norm  # noqa: E402  (reference)
from vit_pytorch_face import ViT_face  # noqa: E402  (reference)
import loralib as lora  # noqa: E402

from oracle.vit_oracle import TINY, VitConfig, init_state_dict  # noqa: E402


def main():
    assert ref_engine.__file__.startswith(REF), ref_engine.__file__
    gold = {}
    for depth in (6, 3):
        cfg = VitConfig(**{**TINY.to_dict(), "depth": depth})
        seed = 40 + depth
        sd = init_state_dict(cfg, seed=seed)
        m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                     depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                     lora_rank=cfg.lora_rank)
        m.load_state_dict(sd, strict=True)
        lora.mark_only_lora_as_trainable(m)
        rec = dict(cfg=cfg.to_dict(), seed=seed, structure={}, norms={})
        for gt in ("block", "lora", "matrix"):
            rec["structure"][gt] = float(ref_engine.get_structure_loss(m, num_layers=depth, group_type=gt, group_pos="FFN"))
            for typ in ("L2", "L1"):
                vals = cal_norm.get_norm_of_lora(m, type=typ, group_num=depth, group_type=gt, group_pos="FFN")
                rec["norms"][f"{gt}_{typ}"] = [float(v) for v in vals]
        gold[f"depth{depth}"] = rec
        print(depth, rec["structure"], {k: len(v) for k, v in rec["norms"].items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_groupings.pt")
    torch.save(gold, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
