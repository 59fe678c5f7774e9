"""gslora-b200 `engine` -- drop-in for the reference's single-step forgetting loop (engine.py of bjzhb666/GS-LoRA, imported by
train/train_own_forget.py:40 as `from engine import train_one_epoch, eval_data`): `train_one_epoch` (engine.py:13-434), `evaluate` (:436),
`eval_data` (:501), `get_structure_loss(model, num_layers, group_type, group_pos)` (:532-687) and `get_prototype_loss` (:690).

Same fused step as engine_cl (one engine forward over remain + forget, device-side losses and gates, selective backward, group-Lasso
AdamW); what differs from engine_cl is host-side and is reproduced here:
  * few-shot loader swap (engine.py:53-57): when the forget loader is the longer one and cfg["few_shot"] is set, the FORGET loader drives
    the epoch and the remain loader is the prefetched, recycled one;
  * cfg["ALPHA_EPOCH"] gating of the structure term (engine.py:82-90), cfg["GROUP_TYPE"] in {"block", "lora", "matrix"};
  * the prototype bound is the literal 18 (engine.py:104);
  * evaluation runs on a deep copy in the reference (engine.py:449, 513), i.e. the model being trained is never merged: here the engine
    evaluates the live (un-merged) LoRA weights with dropout off, which is the same function without copying 19 M parameters.
"""
from __future__ import annotations

import os

import torch

import engine_cl as _cl
from engine_cl import AverageMeter, StepResult, get_prototype_loss, get_time, unlearn_step, unlearn_step_async  # noqa: F401
from gslora import _ffi as F

_PROTO_BOUND = 18.0         # engine.py:104


def get_structure_loss(model: torch.nn.Module, num_layers: int = None, group_type: str = "block", group_pos: str = "FFN"):
    """engine.get_structure_loss (engine.py:532-687).  `num_layers` is implied by the engine-backed model (the reference needs it only to
    build parameter names)."""
    m = _cl._unwrap(model)
    _check_group_pos(m, group_pos)
    if group_type not in ("block", "lora", "matrix"):
        raise ValueError(f"group_type {group_type!r} not in block / lora / matrix")
    if num_layers is not None and int(num_layers) != m.engine_spec().depth:
        raise ValueError(f"num_layers={num_layers} but the model has {m.engine_spec().depth} Transformer blocks")
    return _cl.get_structure_loss(model, group_type=group_type)


def _check_group_pos(m, group_pos):
    """group_pos "FFN" / "Attention" (engine.py:585-658) names where the LoRA parameters live: it has to agree with how the model was built
    (ViT_face(lora_pos=...)).  The reference, handed the other one, silently sums an empty list of groups."""
    if group_pos not in ("FFN", "Attention"):
        raise ValueError(f"group_pos {group_pos!r} not in FFN / Attention")
    have = getattr(m, "lora_pos", "FFN")
    if have != group_pos:
        raise ValueError(f"group_pos={group_pos!r} but the model carries its LoRA on {have!r} (ViT_face(lora_pos=...))")


def train_one_epoch(model, dataloader_forget, dataloader_remain, device, criterion, optimizer, epoch, losses_forget, losses_remain,
                    losses_total, losses_structure, top1_forget, top1_remain, beta, alpha, BND, batch, testloader_forget, testloader_remain,
                    forget_acc_before, highest_H_mean, cfg, dataloader_open=None, prototype_weight_forget=0.0, prototype_weight_remain=0.0,
                    use_prototype=False, prototype_dict=None, losses_prototype_forget=None, losses_prototype_remain=None):
    """Same contract as engine.train_one_epoch (engine.py:13-434); returns the 10-tuple of engine.py:424-434."""
    model.train()
    criterion.train()
    m = _cl._unwrap(model)
    if _cl.engine_fresh_optimizer(m, optimizer):
        m.ensure_engine(1)
        m._engine.reset_optimizer()
    if losses_prototype_forget is None:
        losses_prototype_forget = AverageMeter()
    if losses_prototype_remain is None:
        losses_prototype_remain = AverageMeter()
    DISP_FREQ, VER_FREQ = 5, 100
    alpha_eff = 0.0 if epoch < cfg.get("ALPHA_EPOCH", 0) else alpha               # engine.py:82-90
    group_type = cfg.get("GROUP_TYPE", "block")
    _check_group_pos(m, cfg.get("GROUP_POS", "FFN"))
    forget_drives = len(dataloader_forget) > len(dataloader_remain) and bool(cfg.get("few_shot"))      # engine.py:53
    driving, recycled = (dataloader_forget, dataloader_remain) if forget_drives else (dataloader_remain, dataloader_forget)
    prefetcher = _cl._Prefetcher(recycled, device)
    side_x, side_y = prefetcher.next()
    rank0 = _cl._dist() is None or _cl._dist().get_rank() == 0
    pending = None

    def absorb(p):
        out, nr, nf = p[0].wait(), p[1], p[2]
        losses_remain.update(out["loss_remain"], nr)
        top1_remain.update(out["top1_remain"], nr)
        losses_forget.update(beta * out["loss_forget"], nf)
        top1_forget.update(out["top1_forget"], nf)
        losses_structure.update(alpha_eff * out["structure"], nr)
        losses_prototype_forget.update(prototype_weight_forget * max(_PROTO_BOUND - out["proto_forget"], 0.0), nr)
        losses_prototype_remain.update(out["proto_remain"] * prototype_weight_remain, nr)
        losses_total.update(out["total"], nr)

    for main_x, main_y in iter(driving):
        n_main = int(main_x.size(0))                            # the meters weigh by the GLOBAL batch sizes
        main_x, main_y = _cl.shard_batch(main_x, main_y)        # sharded on the host: each rank copies only its share over PCIe
        main_x, main_y = main_x.to(device), main_y.to(device)
        (side_x, side_y), n_side = _cl._own_share(prefetcher, side_x, side_y)
        (xr, yr), (xf, yf) = ((side_x, side_y), (main_x, main_y)) if forget_drives else ((main_x, main_y), (side_x, side_y))
        n_r, n_f = (n_side, n_main) if forget_drives else (n_main, n_side)
        res = unlearn_step_async(model, xr, yr, xf, yf, beta=beta, alpha=alpha_eff, BND=BND, optimizer=optimizer,
                                 use_prototype=use_prototype, prototype_dict=prototype_dict, prototype_weight_forget=prototype_weight_forget,
                                 prototype_weight_remain=prototype_weight_remain, BND_pro=_PROTO_BOUND if use_prototype else 0.0,
                                 group_type=group_type)
        if pending is not None:
            absorb(pending)
        pending = (res, n_r, n_f)
        show = ((batch + 1) % DISP_FREQ == 0) and batch != 0
        verify = ((batch + 1) % VER_FREQ == 0) and batch != 0
        if show or verify:
            absorb(pending)
            pending = None
        if show:
            if rank0:
                _cl._wandb_log({"epoch_loss_forget": losses_forget.avg, "epoch_loss_remain": losses_remain.avg,
                                "epoch_acc_forget": top1_forget.avg, "epoch_acc_remain": top1_remain.avg,
                                "epoch_loss_total": losses_total.avg, "epoch_loss_structure": losses_structure.avg,
                                "epoch_loss_prototype_forget": losses_prototype_forget.avg,
                                "epoch_loss_prototype_remain": losses_prototype_remain.avg})
                print("Epoch {} Batch {}\t"
                      "Training forget Loss {lf.val:.4f} ({lf.avg:.4f})\tTraining remain Loss {lr.val:.4f} ({lr.avg:.4f})\t"
                      "Training forget prototype Loss {pf.val:.4f}\tTraining remain prototype Loss {pr.val:.4f}\t"
                      "Training structure Loss {ls.val:.4f} ({ls.avg:.4f})\tTraining total Loss {lt.val:.4f} ({lt.avg:.4f})\t"
                      "Training forget Prec@1 {tf.val:.3f} ({tf.avg:.3f})\tTraining remain Prec@1 {tr.val:.3f} ({tr.avg:.3f})".format(
                          epoch + 1, batch + 1, lf=losses_forget, lr=losses_remain, pf=losses_prototype_forget, pr=losses_prototype_remain,
                          ls=losses_structure, lt=losses_total, tf=top1_forget, tr=top1_remain))
            losses_forget, losses_remain, top1_forget, top1_remain = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
            losses_total, losses_structure = AverageMeter(), AverageMeter()
            losses_prototype_forget, losses_prototype_remain = AverageMeter(), AverageMeter()
        if verify:
            with torch.no_grad():
                highest_H_mean = evaluate(model, testloader_forget=testloader_forget, testloader_remain=testloader_remain, device=device,
                                          batch=batch, epoch=epoch, forget_acc_before=forget_acc_before, highest_H_mean=highest_H_mean,
                                          cfg=cfg, optimizer=optimizer, testloader_open=dataloader_open)
        batch += 1
        side_x, side_y = prefetcher.next()
        if side_x is None:
            prefetcher = _cl._Prefetcher(recycled, device)
            side_x, side_y = prefetcher.next()
    if pending is not None:
        absorb(pending)
    _cl.sync_optimizer_state(model, optimizer)
    return (batch, highest_H_mean, losses_forget, losses_remain, top1_forget, top1_remain, losses_total, losses_structure,
            losses_prototype_forget, losses_prototype_remain)


def evaluate(model, testloader_forget, testloader_remain, device, batch, epoch, forget_acc_before, highest_H_mean, cfg, optimizer,
             testloader_open=None):
    """engine.evaluate (engine.py:436-498): accuracies of the current weights, H-mean, rolling best checkpoint (at most 2 kept)."""
    lr = optimizer.param_groups[0]["lr"]
    print("current learning rate:{:.7f}".format(lr))
    print("Perfom evaluation on test set and save checkpoints...")
    forget_acc = eval_data(model, testloader_forget, device, "forget", batch)
    remain_acc = eval_data(model, testloader_remain, device, "remain", batch)
    if testloader_open is not None:
        eval_data(model, testloader_open, device, "open", batch)
    forget_drop = forget_acc_before - forget_acc
    denom = forget_drop + remain_acc
    Hmean = 2 * forget_drop * remain_acc / denom if denom != 0 else 0.0      # the reference divides unguarded (engine.py:460)
    rank0 = _cl._dist() is None or _cl._dist().get_rank() == 0
    if Hmean > highest_H_mean:
        highest_H_mean = Hmean
        if rank0:
            path = os.path.join(cfg["WORK_PATH"], "Backbone_{}_Epoch_{}_Batch_{}_Time_{}_checkpoint.pth".format(
                cfg["BACKBONE_NAME"], epoch + 1, batch + 1, get_time()))
            torch.save(_merged_state_dict(_cl._unwrap(model)), path)      # the reference saves the eval()-ed deep copy: merged weights
            if len(os.listdir(cfg["WORK_PATH"])) >= 3:
                ckpts = sorted((f for f in os.listdir(cfg["WORK_PATH"]) if f.endswith(".pth")),
                               key=lambda f: os.path.getmtime(os.path.join(cfg["WORK_PATH"], f)))
                os.remove(os.path.join(cfg["WORK_PATH"], ckpts[0]))
    return highest_H_mean


def _merged_state_dict(m):
    """state_dict of `copy.deepcopy(model).eval()` without the copy: loralib's merge (W + B A * scaling, loralib Linear.train(False)) applied to
    the weights of every un-merged LoRA layer in a cloned dict; the training model itself stays un-merged."""
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for name, mod in m.named_modules():
        if getattr(mod, "r", 0) > 0 and hasattr(mod, "lora_A") and hasattr(mod, "lora_B") and not getattr(mod, "merged", False):
            sd[(name + "." if name else "") + "weight"] += (mod.lora_B.detach() @ mod.lora_A.detach()) * mod.scaling
    return sd


def eval_data(model, dataloader, device, mode: str, batch: int = 0):
    """engine.eval_data (engine.py:501-529): top-1 accuracy (0-100) of the current weights in eval mode.  The reference evaluates a deep copy
    (so the training model keeps its mode and stays un-merged); the engine runs the live LoRA weights with dropout off instead."""
    m = _cl._unwrap(model)
    hits = torch.zeros((), dtype=torch.int64, device=device)
    total = 0
    with torch.no_grad():
        for images, labels in dataloader:
            images = m.prepare_images(images.to(device))
            labels = labels.to(device).long().contiguous()
            for slot, _, B in m.inference_slots(images, labels):          # chunks of the engine's capacity: eval never grows the workspace
                hits += m._engine.slot_tensor(slot, F.SLOT_CORRECT, B).sum()
            total += labels.size(0)
    accuracy = 100 * int(hits.item()) / max(total, 1)
    print("Test {} Accuracy:{:2f}%".format(mode, accuracy))
    _cl._wandb_log({"Test {} Accuracy".format(mode): accuracy})
    return accuracy
