"""Restatement of timm==0.9.2 create_optimizer for opt == 'adamw' (reference call site
train/train_own_forget_cl.py:811-813): torch.optim.AdamW over requires_grad params, 1-D / .bias
params in a weight_decay=0 group, the rest decayed by args.weight_decay.  PARITY UNPINNED (no
reference test at this boundary)."""
import torch


def create_optimizer(args, model, filter_bias_and_bn=True):
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.ndim <= 1 or name.endswith(".bias"):
            no_decay.append(p)
        else:
            decay.append(p)
    wd = getattr(args, "weight_decay", 0.05)
    groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd}]
    kw = dict(lr=args.lr, eps=getattr(args, "opt_eps", None) or 1e-8)
    betas = getattr(args, "opt_betas", None)
    if betas:
        kw["betas"] = tuple(betas)
    return torch.optim.AdamW(groups, weight_decay=0.0, **kw)
