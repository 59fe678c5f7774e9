"""Shared by the free-running trajectory tests: the scenario of tests/golden/make_golden_trajectory.py (loralib's start lora_B = 0, lr 1e-2,
ALPHA_EPOCH switch 0 -> alpha = 2 after 15 steps, a forget bound the run reaches) replayed step by step on the oracle."""
import torch

from oracle import vit_oracle as O


COLLAPSED = 0.12     # a group counts as collapsed when its norm fell below 12 % of its norm at the ALPHA_EPOCH switch


def trajectory_loaders(cfg, seed, n, bs, distinct=2):
    """identical to make_golden_trajectory.trajectory_loaders (kept here so the GPU box, which has no reference tree, can rebuild the batches)"""
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    rb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    fb = [(torch.rand(bs, 3, S, S, generator=g), torch.randint(0, cfg.num_class, (bs,), generator=g)) for _ in range(distinct)]
    return [rb[i % distinct] for i in range(n)], [fb[i % distinct] for i in range(n)]


def group_norms(params, cfg):
    return [float(torch.sqrt(sum((params[n].double() ** 2).sum() for n in grp))) for grp in O.lora_names(cfg)]


def oracle_trajectory(cfg, sd, hp, device="cpu", loaders=None, grad_noise=0.0, noise_seed=0):
    """Free-running oracle: per-step dicts (total, ce_forget, loss_remain, structure) and the group norms after each epoch.
    grad_noise = eps adds eps * rms(g) * N(0, 1) to every LoRA gradient element before AdamW: the FP32 reference perturbed at a given relative
    gradient tolerance (rel-L2 error eps per tensor) -- the envelope any implementation that meets that tolerance step by step lives in.  Adam
    normalises every element's update to ~lr whatever its gradient's size, so elements whose true gradient is far below eps * rms(g) turn
    from a consistent drift into a random walk: the trajectory is far more sensitive to ADDITIVE error than the 1e-3 suggests."""
    sd = {k: v.clone().to(device) for k, v in sd.items()}
    state, steps, norms = {}, [], []
    gen = torch.Generator(device=device).manual_seed(1000 + noise_seed)
    orig = O.unlearn_grads

    def noisy(*a, **k):
        out, gd = orig(*a, **k)
        return out, {n: v + grad_noise * v.pow(2).mean().sqrt() * torch.randn(v.shape, generator=gen, device=v.device) for n, v in gd.items()}
    if grad_noise:
        O.unlearn_grads = noisy
    try:
        return _oracle_trajectory(cfg, sd, hp, device, loaders, state, steps, norms)
    finally:
        O.unlearn_grads = orig


def _oracle_trajectory(cfg, sd, hp, device, loaders, state, steps, norms):
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


def deviation(steps, norms, ref_steps, ref_norms, hp):
    """How far a trajectory is from the reference one: worst 5-step-window total loss, worst per-step total loss, worst final group norm (all
    relative), the set of collapsed groups and the step at which CE_forget first reaches BND."""
    w, wr = windows(steps, hp), windows(ref_steps, hp)
    return dict(window_total=max(abs(a["epoch_loss_total"] - b["epoch_loss_total"]) / abs(b["epoch_loss_total"]) for a, b in zip(w, wr)),
                step_total=max(abs(a["total"] - b["total"]) / abs(b["total"]) for a, b in zip(steps, ref_steps)),
                final_norm=max(abs(a - b) / b for a, b in zip(norms[-1], ref_norms[-1])),
                collapsed=[n1 < COLLAPSED * n0 for n0, n1 in zip(norms[0], norms[-1])],
                first_cross=next((i for i, s_ in enumerate(steps) if s_["ce_forget"] >= hp["BND"]), None))


def noise_envelope(cfg, sd, hp, eps, device, loaders=None, seeds=(0, 1, 2), ref=None):
    """Per-metric maximum deviation of the FP32 oracle from itself under additive gradient noise eps (see oracle_trajectory)."""
    ref_steps, ref_norms = ref if ref is not None else oracle_trajectory(cfg, sd, hp, device=device, loaders=loaders)[:2]
    devs = []
    for k in seeds:
        st, nm, _ = oracle_trajectory(cfg, sd, hp, device=device, loaders=loaders, grad_noise=eps, noise_seed=k)
        devs.append(deviation(st, nm, ref_steps, ref_norms, hp))
    env = {m: max(d[m] for d in devs) for m in ("window_total", "step_total", "final_norm")}
    env["first_cross"] = sorted(d["first_cross"] for d in devs if d["first_cross"] is not None)
    env["collapsed"] = [d["collapsed"] for d in devs]
    return env
