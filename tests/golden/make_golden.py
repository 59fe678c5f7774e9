"""Generate golden vectors by running the UNMODIFIED reference (bjzhb666/GS-LoRA at /root/reference).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It puts `oracle/shims` (IPython/mxnet/matplotlib/swanlab/timm stand-ins + the loralib 0.1.2
restatement -- loralib itself is absent, see oracle/loralib_restated.py) ahead of /root/reference
on sys.path, imports the reference's own `vit_pytorch_face.ViT_face`, `engine_cl`, `util.cal_norm`,
and records, on seeded synthetic inputs:
  logits / emb for both streams, the step losses, every LoRA gradient of
  `loss_total.backward()` (engine_cl.py:118-124), engine_cl.get_structure_loss,
  engine_cl.get_prototype_loss, util.cal_norm.get_norm_of_lora, and the LoRA tensors after two
  torch.optim.AdamW steps built by the timm-restated create_optimizer.
The only liberty taken: `torch.Tensor.cuda` is patched to the identity so CosFace's unconditional
`label.cuda(self.device_id[0])` (vit_face.py:201) runs on CPU; arithmetic is untouched.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF, ROOT]
os.environ.setdefault("WANDB_MODE", "disabled")

import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self  # CPU stand-in for vit_face.py:176-201 device hops

# image_iter.CustomSubset breaks on torch 2.11 (SURVEY 8b landmine 2); util.utils only needs the name.
_ii = types.ModuleType("image_iter")
_ii.CustomSubset = type("CustomSubset", (), {})
sys.modules["image_iter"] = _ii

import engine_cl  # noqa: E402  (reference)
import util.cal_norm as cal_norm  # noqa: E402  (reference)
from vit_pytorch_face import ViT_face  # noqa: E402  (reference)
import loralib as lora  # noqa: E402  (oracle restatement via shim)
from timm.optim import create_optimizer  # noqa: E402

from oracle.vit_oracle import VitConfig, TINY, P8S8, init_state_dict, lora_param_list  # noqa: E402


def build_reference_model(cfg: VitConfig, sd):
    torch.manual_seed(0)
    m = ViT_face(loss_type="CosFace", GPU_ID=[0], num_class=cfg.num_class, image_size=cfg.image_size,
                 patch_size=cfg.patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
                 dim_head=cfg.dim_head, dropout=0.0, emb_dropout=0.0, lora_rank=cfg.lora_rank)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    lora.mark_only_lora_as_trainable(m)
    m.train()
    return m


def run_case(cfg: VitConfig, seed: int, B: int, hp: dict, steps: int, keep_weights: bool):
    sd = init_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    S = cfg.image_size
    img_r = torch.rand(B, 3, S, S, generator=g)
    img_f = torch.rand(B, 3, S, S, generator=g)
    lab_r = torch.randint(0, cfg.num_class, (B,), generator=g)
    lab_f = torch.randint(0, cfg.num_class, (B,), generator=g)
    protos = torch.randn(cfg.num_class, cfg.dim, generator=g)
    proto_dict = {i: protos[i] for i in range(cfg.num_class)}

    model = build_reference_model(cfg, sd)
    crit = torch.nn.CrossEntropyLoss()
    args = types.SimpleNamespace(lr=hp["lr"], weight_decay=hp["wd"], opt_eps=1e-8, opt_betas=None, opt="adamw")
    opt = create_optimizer(args, model)

    names = lora_param_list(cfg)
    gold = dict(cfg=cfg.to_dict(), seed=seed, hp=hp, B=B, img_r=img_r, img_f=img_f, lab_r=lab_r, lab_f=lab_f,
                prototypes=protos)
    if keep_weights:
        gold["state_dict"] = {k: v.clone() for k, v in sd.items()}
    # weights are otherwise regenerated from the seed by oracle.vit_oracle.init_state_dict; pin them
    gold["state_dict_checksum"] = {k: float(v.double().abs().sum()) for k, v in sd.items()}
    per_step = []
    for step in range(steps):
        # engine_cl.py:59-125, dropout 0, use_prototype per hp
        out_r, emb_r = model(img_r.float(), lab_r)
        loss_remain = crit(out_r, lab_r)
        out_f, emb_f = model(img_f.float(), lab_f)
        ce_f = crit(out_f, lab_f)
        loss_forget = torch.functional.F.relu(hp["BND"] - ce_f)
        s_loss = engine_cl.get_structure_loss(model)
        if hp.get("use_proto"):
            pf = engine_cl.get_prototype_loss(emb_f, lab_f, proto_dict)
            pr = engine_cl.get_prototype_loss(emb_r, lab_r, proto_dict)
            proto = hp["w_pf"] * torch.functional.F.relu(hp["BND_pro"] - pf) + hp["w_pr"] * pr
        else:
            pf = pr = proto = torch.tensor(0.0)
        total = loss_forget * hp["beta"] + loss_remain + s_loss * hp["alpha"] + proto
        opt.zero_grad()
        total.backward()
        rec = dict(logits_r=out_r.detach().clone(), logits_f=out_f.detach().clone(), emb_r=emb_r.detach().clone(),
                   emb_f=emb_f.detach().clone(), loss_remain=loss_remain.item(), ce_forget=ce_f.item(),
                   loss_forget=loss_forget.item(), structure=s_loss.item(), proto_forget=float(pf),
                   proto_remain=float(pr), total=total.item(),
                   grads={n: model.get_parameter(n).grad.detach().clone() for n in names})
        opt.step()
        rec["params_after"] = {n: model.get_parameter(n).detach().clone() for n in names}
        per_step.append(rec)
    gold["steps"] = per_step
    gold["norm_of_lora_L2"] = [float(x) for x in cal_norm.get_norm_of_lora(model, type="L2", group_num=cfg.depth)]
    gold["norm_of_lora_L1"] = [float(x) for x in cal_norm.get_norm_of_lora(model, type="L1", group_num=cfg.depth)]
    # eval-mode (merged) forward: loralib merge semantics, engine_cl.eval_data path
    model.eval()
    with torch.no_grad():
        out_e, emb_e = model(img_r.float(), lab_r)
    gold["eval_logits_r"] = out_e.clone()
    gold["eval_merged_fc1_w0"] = model.get_parameter("transformer.layers.0.1.fn.fn.net.0.weight").detach()[:4, :8].clone()
    model.train()
    return gold


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    hp = dict(lr=1e-2, wd=0.05, beta=0.15, alpha=1e-4, BND=105.0)
    hp_proto = dict(hp, use_proto=True, w_pf=1.0, w_pr=1.0, BND_pro=18.0)
    # depth-6 TINY-width variant so engine_cl.get_structure_loss's hard-coded 6 groups resolve
    tiny6 = VitConfig(**{**TINY.to_dict(), "depth": 6})
    cases = {
        "tiny6_b4": (tiny6, 11, 4, hp, 2, True),
        "tiny6_b4_proto": (tiny6, 12, 4, hp_proto, 1, False),
        "tiny6_b3_lowbnd": (tiny6, 13, 3, dict(hp, BND=2.0), 1, False),   # gate closed: relu(BND-CE)=0
    }
    for name, (cfg, seed, B, h, steps, keep) in cases.items():
        gold = run_case(cfg, seed, B, h, steps, keep)
        path = os.path.join(out_dir, f"{name}.pt")
        torch.save(gold, path)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "total", gold["steps"][0]["total"])
    # P8S8 (config 2 shape) at B=2: weights are regenerated from the seed (too large to commit);
    # keep logits/emb/losses and a strided subsample of every gradient tensor.
    gold = run_case(P8S8, 1337, 2, hp, 1, False)
    for rec in gold["steps"]:
        rec["grad_norms"] = {n: float(g.norm()) for n, g in rec["grads"].items()}
        rec["grads"] = {n: g.flatten()[::37].clone() for n, g in rec["grads"].items()}
        rec["params_after"] = {n: p.flatten()[::37].clone() for n, p in rec["params_after"].items()}
    path = os.path.join(out_dir, "p8s8_b2.pt")
    torch.save(gold, path)
    print("p8s8_b2 ->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
