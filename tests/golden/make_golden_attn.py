"""Golden for LoRA on attention: the UNMODIFIED reference `ViT_face(lora_pos="Attention")` (vit_face.py:349-355, 405-425:
lora.MergedLinear(r, enable_lora=[True] * 3) on to_qkv, plain Linears in the FFN) with `engine.get_structure_loss(group_pos="Attention")`
(engine.py:650-656) and `util.cal_norm.get_norm_of_lora(group_pos="Attention")`, on seeded synthetic inputs: logits / emb, the step losses,
every LoRA gradient, parameters after two timm-built AdamW steps, the norm report and the eval-mode (merged) forward.
Same liberties as make_golden.py (Tensor.cuda -> identity; loralib is the oracle restatement via oracle/shims).

    python tests/golden/make_golden_attn.py        # authoring container only (needs /root/reference)
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
from timm.optim import create_optimizer  # noqa: E402

from oracle.vit_oracle import TINY, VitConfig, init_state_dict, lora_param_list  # noqa: E402


def run_case(cfg, seed, B, hp, steps):
    sd = init_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    S = cfg.image_size
    img_r, img_f = torch.rand(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g)
    lab_r, lab_f = torch.randint(0, cfg.num_class, (B,), generator=g), torch.randint(0, cfg.num_class, (B,), generator=g)
    torch.manual_seed(0)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size, patch_size=cfg.patch_size, dim=cfg.dim,
                 depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim, dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0,
                 lora_rank=cfg.lora_rank, lora_pos="Attention")
    m.load_state_dict(sd, strict=True)               # key set: to_qkv.lora_A / lora_B present, no FFN lora_* keys
    lora.mark_only_lora_as_trainable(m)
    m.train()
    trainable = sorted(n for n, p in m.named_parameters() if p.requires_grad)
    assert trainable == sorted(lora_param_list(cfg)), trainable
    crit = torch.nn.CrossEntropyLoss()
    opt = create_optimizer(types.SimpleNamespace(lr=hp["lr"], weight_decay=hp["wd"], opt_eps=1e-8, opt_betas=None, opt="adamw"), m)
    names = lora_param_list(cfg)
    gold = dict(cfg=cfg.to_dict(), seed=seed, hp=hp, B=B, img_r=img_r, img_f=img_f, lab_r=lab_r, lab_f=lab_f,
                state_dict={k: v.clone() for k, v in sd.items()})
    recs = []
    for _ in range(steps):
        out_r, emb_r = m(img_r.float(), lab_r)
        loss_remain = crit(out_r, lab_r)
        out_f, emb_f = m(img_f.float(), lab_f)
        ce_f = crit(out_f, lab_f)
        loss_forget = torch.functional.F.relu(hp["BND"] - ce_f)
        s_loss = ref_engine.get_structure_loss(m, num_layers=cfg.depth, group_type="block", group_pos="Attention")
        total = loss_forget * hp["beta"] + loss_remain + s_loss * hp["alpha"]
        opt.zero_grad()
        total.backward()
        rec = dict(logits_r=out_r.detach().clone(), logits_f=out_f.detach().clone(), emb_r=emb_r.detach().clone(), loss_remain=loss_remain.item(),
                   ce_forget=ce_f.item(), loss_forget=loss_forget.item(), structure=s_loss.item(), total=total.item(),
                   grads={n: m.get_parameter(n).grad.detach().clone() for n in names})
        opt.step()
        rec["params_after"] = {n: m.get_parameter(n).detach().clone() for n in names}
        recs.append(rec)
    gold["steps"] = recs
    gold["norm_of_lora_L2"] = [float(x) for x in cal_norm.get_norm_of_lora(m, type="L2", group_num=cfg.depth, group_pos="Attention")]
    gold["norm_of_lora_L1"] = [float(x) for x in cal_norm.get_norm_of_lora(m, type="L1", group_num=cfg.depth, group_pos="Attention")]
    m.eval()
    with torch.no_grad():
        out_e, _ = m(img_r.float(), lab_r)
    gold["eval_logits_r"] = out_e.clone()
    gold["eval_merged_qkv_w"] = m.get_parameter("transformer.layers.0.0.fn.fn.to_qkv.weight").detach()[::37, :16].clone()
    m.train()
    return gold


def main():
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=1e-2, BND=105.0)
    cfg = VitConfig(**{**TINY.to_dict(), "depth": 3, "lora_pos": "Attention"})
    gold = run_case(cfg, 21, 4, hp, 2)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny3_attn_lora.pt")
    torch.save(gold, path)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB", "total", gold["steps"][0]["total"], "structure", gold["steps"][0]["structure"],
          "norms", gold["norm_of_lora_L2"])


if __name__ == "__main__":
    main()
