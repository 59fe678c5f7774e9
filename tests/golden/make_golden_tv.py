"""Golden vector for the torchvision family (BASELINE configs 4 / 5): the UNMODIFIED reference wrapper `vit_pytorch_face.ModifiedViT`
(modified_VIT.py:5-39) around a torchvision VisionTransformer whose MLP Linears were swapped by the UNMODIFIED
`util.utils.replace_ffn_with_lora` (util/utils.py:552-576); logits, cls embedding, CE loss and every LoRA gradient on a seeded batch.

    python tests/golden/make_golden_tv.py        # authoring container only (needs /root/reference)
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF, ROOT]
os.environ.setdefault("WANDB_MODE", "disabled")

import torch  # noqa: E402

_ii = types.ModuleType("image_iter")
_ii.CustomSubset = type("CustomSubset", (), {})
sys.modules["image_iter"] = _ii

import util.utils as ref_utils  # noqa: E402  (reference)
import vit_pytorch_face.modified_VIT as ref_mv  # noqa: E402  (reference)
import loralib as lora  # noqa: E402
from torchvision.models.vision_transformer import VisionTransformer  # noqa: E402

SHAPE = dict(image_size=64, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)
RANK, SEED, B = 8, 31, 3


def main():
    assert ref_utils.__file__.startswith(REF) and ref_mv.__file__.startswith(REF)
    torch.manual_seed(SEED)
    model = ref_mv.ModifiedViT(VisionTransformer(**SHAPE))
    ref_utils.replace_ffn_with_lora(model, rank=RANK)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "lora_B" in n or n.endswith("heads.head.weight"):
                p.normal_(0, 0.02)              # fresh lora_B and torchvision's head are zero: make both live
        model.encoder.pos_embedding.normal_(0, 0.02)
        model.class_token.normal_(0, 0.02)
    lora.mark_only_lora_as_trainable(model)
    model.train()
    g = torch.Generator().manual_seed(SEED + 1)
    x = torch.randn(B, 3, SHAPE["image_size"], SHAPE["image_size"], generator=g)
    y = torch.randint(0, SHAPE["num_classes"], (B,), generator=g)
    logits, emb = model(x, y)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    gold = dict(shape=SHAPE, rank=RANK, x=x, y=y, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                logits=logits.detach().clone(), emb=emb.detach().clone(), loss=float(loss), trainable=names,
                grads={n: model.get_parameter(n).grad.detach().clone() for n in names})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tv_small_b3.pt")
    torch.save(gold, path)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB", "loss", float(loss), len(names), "trainable tensors")


if __name__ == "__main__":
    main()
