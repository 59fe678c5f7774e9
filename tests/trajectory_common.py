"""Shared by the free-running trajectory tests: the scenario of tests/golden/make_golden_trajectory.py (loralib's start lora_B = 0, lr 1e-2,
ALPHA_EPOCH switch 0 -> alpha = 2 after 15 steps, a forget bound the run reaches) replayed step by step on the oracle."""
import torch

from oracle import vit_oracle as O


def trajectory_loaders(cfg, seed, n, bs, distinct=2):
    """identical to make_golden_trajectory.trajectory_loaders (kept here so the GPU box, which has no reference tree, can rebuild the batches)"""
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    rb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    fb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    return [rb[i % distinct] for i in range(n)], [fb[i % distinct] for i in range(n)]


def group_norms(params, cfg):
    return [float(torch.sqrt(sum((params[n].double() ** 2).sum() for n in grp))) for grp in O.lora_names(cfg)]


def oracle_trajectory(cfg, sd, hp, device="cpu", loaders=None):
    """Free-running oracle: per-step dicts (total, ce_forget, loss_remain, structure) and the group norms after each epoch."""
    sd = {k: v.clone().to(device) for k, v in sd.items()}
    state, steps, norms = {}, [], []
    for epoch, n in enumerate(hp["steps"]):
        remain, forget = loaders(epoch, n) if loaders else trajectory_loaders(cfg, hp["seed"] + 10, n, hp["batch"])
        alpha = 0.0 if epoch < hp["alpha_epoch"] else hp["alpha"]
        for (xr, yr), (xf, yf) in zip(remain, forget):
            out, _ = O.unlearn_step(sd, cfg, state, xr.to(device), yr.to(device), xf.to(device), yf.to(device), lr=hp["lr"], wd=hp["wd"],
                                    beta=hp["beta"], alpha=alpha, BND=hp["BND"], include_structure=alpha != 0.0)
            steps.append(dict(total=float(out["total"]), ce_forget=float(out["ce_forget"]), loss_remain=float(out["loss_remain"]),
                              loss_forget=float(out["loss_forget"]), structure=float(out["structure"]), alpha=alpha))
        norms.append(group_norms({n_: sd[n_] for n_ in O.lora_param_list(cfg)}, cfg))
    return steps, norms, sd


def windows(steps, hp, width=5):
    """The reference loop's display records (engine.py:128-187): meters averaged over each 5-step window, reset after every display."""
    out = []
    for lo in range(0, len(steps), width):
        w = steps[lo:lo + width]
        out.append(dict(epoch_loss_forget=sum(hp["beta"] * s["loss_forget"] for s in w) / len(w),
                        epoch_loss_remain=sum(s["loss_remain"] for s in w) / len(w),
                        epoch_loss_total=sum(s["total"] for s in w) / len(w),
                        epoch_loss_structure=sum(s["alpha"] * s["structure"] for s in w) / len(w)))
    return out
