"""Replay of the call sequence of the UNMODIFIED continual-forgetting driver (train/train_own_forget_cl.py, `--one_stage`, the GS-LoRA path)
against the drop-in surface, for the GPU box -- where the reference tree does not exist, so the driver itself cannot be launched.  Every stage
cites the driver lines it re-enacts; only host orchestration is restated here (dataset splits are synthetic tensors), every model call goes
through the same public entry points the driver imports: ViT_face / ModifiedViT, loralib, engine_cl.train_one_epoch / eval_data,
util.cal_norm.get_norm_of_lora, gslora.prototypes.calculate_prototypes.

tests/test_driver_dropin_cpu.py runs the real driver in the authoring container up to its first engine call; this harness covers what comes after
it: eval x4 (+ old), the epoch loop with per-epoch cosine lr, EMA deep copies, the norm report, eval-mode (merged-weight) task checkpoints,
reload + LoRA re-initialisation + a fresh optimizer for the next task."""
import copy
import math
import os

import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset


def reinitialize_lora_parameters(model):
    """util/utils.py:428-441, through the drop-in `util.utils` (the reference's own function when its tree is importable, the overlay's
    restatement on the GPU box)"""
    from util.utils import reinitialize_lora_parameters as f
    return f(model)


def timm_adamw(model, lr, weight_decay):
    """timm.optim.create_optimizer(args, model) for opt='adamw' (train_own_forget_cl.py:811-813): requires_grad parameters, no decay on 1-D / bias"""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if p.requires_grad:
            (no_decay if p.ndim <= 1 or n.endswith(".bias") else decay).append(p)
    return torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}], lr=lr, eps=1e-8)


def cosine_lr(opt, epoch, epochs, lr, min_lr):
    """timm CosineLRScheduler.step(epoch) with warm-up 0 (train_own_forget_cl.py:1013)"""
    v = min_lr + 0.5 * (lr - min_lr) * (1 + math.cos(math.pi * epoch / epochs))
    for g in opt.param_groups:
        g["lr"] = v
    return v


def replay(backbone, *, image_size, num_class, device, work_path, num_tasks=2, epochs=1, batch_size=8, first_cls=None, per_forget=None,
           per_class=4, lr=1e-2, min_lr=1e-5, wd=0.05, BND=105.0, alpha=1e-4, betas=(0.3, 0.4, 0.28, 0.2), prototype=True, average_weight=True,
           ema_epoch=0, ema_decay=0.9, imagenet=False, seed=0, log=None):
    import engine_cl
    from gslora.prototypes import calculate_prototypes
    from util.cal_norm import get_norm_of_lora
    log = log if log is not None else []
    g = torch.Generator().manual_seed(seed)
    first_cls = first_cls if first_cls is not None else num_class - 2 * (per_forget or 2)
    per_forget = per_forget or 2
    # synthetic "dataset": per_class train + 2 test images per class (the driver's ImageFolder + CustomSubset splits, :545-760)
    def make(n):
        y = torch.arange(num_class).repeat_interleave(n)
        return torch.rand(len(y), 3, image_size, image_size, generator=g), y
    xtr, ytr = make(per_class)
    xte, yte = make(2)

    def subset(x, y, lo, hi):
        m = (y >= lo) & (y < hi)
        return TensorDataset(x[m], y[m])

    BACKBONE = backbone.to(device)                                                          # :494-500
    ema_model = None
    if average_weight:                                                                      # :502-507
        BACKBONE.eval()
        ema_model = copy.deepcopy(BACKBONE).to(device)
    BACKBONE.train()
    cfg = {"WORK_PATH": work_path, "BACKBONE_NAME": "VIT", "BND_pro": 18.0, "MULTI_GPU": False}
    os.makedirs(os.path.join(work_path, "task-level"), exist_ok=True)
    out = {"tasks": []}
    for task_i in range(num_tasks):
        if task_i > 0:                                                                      # :524-536
            BACKBONE.load_state_dict(torch.load(os.path.join(work_path, "task-level", f"Backbone_task_{task_i - 1}.pth")))
            reinitialize_lora_parameters(BACKBONE)
            log.append(("reload+reinit", task_i))
        en1 = first_cls - task_i * per_forget                                               # :537-560
        st2, en2 = en1, en1 + per_forget
        forget_tr, remain_tr = subset(xtr, ytr, st2, en2), subset(xtr, ytr, 0, en1)
        forget_te, remain_te = subset(xte, yte, st2, en2), subset(xte, yte, 0, en1)
        old_te = subset(xte, yte, en2, num_class) if task_i > 0 else None
        proto = None
        if prototype:                                                                       # :633-646
            proto = calculate_prototypes(backbone=BACKBONE, dataset=torch.utils.data.ConcatDataset([forget_tr, remain_tr]), device=device, batch_size=500)
            log.append(("prototypes", len(proto)))
        lg = torch.Generator().manual_seed(seed + 1)
        mk = lambda ds, bs, sh: DataLoader(ds, batch_size=bs, shuffle=sh, generator=lg if sh else None, drop_last=False)     # :676-750
        train_forget, train_remain = mk(forget_tr, batch_size, True), mk(remain_tr, batch_size, True)
        train_forget_t, train_remain_t = mk(forget_tr, batch_size * 5, False), mk(remain_tr, batch_size * 5, False)
        test_forget, test_remain = mk(forget_te, batch_size * 5, False), mk(remain_te, batch_size * 5, False)
        LOSS = nn.CrossEntropyLoss()
        OPTIMIZER = timm_adamw(BACKBONE, lr, wd)                                            # :807-820 (a NEW optimizer per task)
        batch = 0
        acc = {}
        for name, loader in (("forget-train", train_forget_t), ("remain-train", train_remain_t), ("forget", test_forget), ("remain", test_remain)):
            acc[name] = engine_cl.eval_data(BACKBONE, loader, device, f"{name}-{task_i}", batch)                       # :899-937
        if old_te is not None:
            acc["old"] = engine_cl.eval_data(BACKBONE, mk(old_te, batch_size * 5, False), device, f"old-{task_i}", batch)
        log.append(("eval_before", task_i, dict(acc)))
        BACKBONE.train()                                                                    # :1000
        highest_H_mean = 0.0
        meters = [engine_cl.AverageMeter() for _ in range(8)]
        lf, lr_m, tf, tr, lt, ls, lpf, lpr = meters
        for epoch in range(epochs):                                                         # :1004-1056
            cosine_lr(OPTIMIZER, epoch, epochs, lr, min_lr)
            (batch, highest_H_mean, lf, lr_m, tf, tr, lt, ls, lpf, lpr) = engine_cl.train_one_epoch(
                model=BACKBONE, dataloader_forget=train_forget, dataloader_remain=train_remain, testloader_forget=test_forget,
                testloader_remain=test_remain, device=device, criterion=LOSS, optimizer=OPTIMIZER, epoch=epoch, batch=batch, losses_forget=lf,
                top1_forget=tf, losses_remain=lr_m, top1_remain=tr, losses_total=lt, losses_structure=ls, beta=betas[task_i % len(betas)], BND=BND,
                forget_acc_before=acc["forget"], highest_H_mean=highest_H_mean, cfg=cfg, alpha=alpha, task_i=task_i, use_prototype=prototype,
                prototype_dict=proto, prototype_weight_forget=0.5, prototype_weight_remain=0.5, losses_prototype_forget=lpf,
                losses_prototype_remain=lpr)
            if average_weight:                                                              # :1058-1098
                with torch.no_grad():
                    COPY = copy.deepcopy(BACKBONE)
                    ema_model.eval()
                    for p, e in zip(COPY.parameters(), ema_model.parameters()):
                        e.data = p.data.detach() if epoch == ema_epoch else e.data.detach() * ema_decay + p.data.detach() * (1 - ema_decay)
                    acc["forget-ema"] = engine_cl.eval_data(ema_model, test_forget, device, f"forget-ema-{task_i}", batch)
                    acc["remain-ema"] = engine_cl.eval_data(ema_model, test_remain, device, f"remain-ema-{task_i}", batch)
        norm_list = get_norm_of_lora(BACKBONE, type="L2", group_num=len(list(BACKBONE.lora_layers())), imagenet=imagenet)   # :1100-1106
        BACKBONE.eval()                                                                     # :1696-1705: eval-mode save (merged weights + lora_*)
        path = os.path.join(work_path, "task-level", f"Backbone_task_{task_i}.pth")
        torch.save(BACKBONE.state_dict(), path)
        with torch.no_grad():
            probe = BACKBONE(xte[:4].to(device), yte[:4].to(device))
        BACKBONE.train()
        if task_i > 0:                                                                      # :1738-1741: old classes once more after the task
            acc["old_after"] = engine_cl.eval_data(BACKBONE, mk(old_te, batch_size * 5, False), device, f"old-{task_i}", batch)
        out["tasks"].append(dict(acc=acc, steps=batch, total=float(lt.avg) if lt.count else None, norms=[float(n) for n in norm_list], ckpt=path,
                                 probe_logits=(probe[0] if isinstance(probe, tuple) else probe).detach().cpu(), opt_state_keys=len(OPTIMIZER.state)))
        log.append(("task_done", task_i, batch))
    out["probe_x"], out["probe_y"] = xte[:4], yte[:4]
    out["model"], out["ema"] = BACKBONE, ema_model
    return out
