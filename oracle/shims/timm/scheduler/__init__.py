"""Restatement of timm==0.9.2 create_scheduler for sched == 'cosine' (call site
train/train_own_forget_cl.py:815-820, stepped per epoch at :1013).  PARITY UNPINNED."""
import math


class CosineLRScheduler:
    def __init__(self, optimizer, t_initial, lr_min=0.0, warmup_t=0, warmup_lr_init=0.0):
        self.optimizer, self.t_initial, self.lr_min = optimizer, t_initial, lr_min
        self.warmup_t, self.warmup_lr_init = warmup_t, warmup_lr_init
        self.base = [g["lr"] for g in optimizer.param_groups]
        if warmup_t > 0:
            for g in optimizer.param_groups:
                g["lr"] = warmup_lr_init

    def step(self, epoch, metric=None):
        for g, base in zip(self.optimizer.param_groups, self.base):
            if epoch < self.warmup_t:
                lr = self.warmup_lr_init + epoch * (base - self.warmup_lr_init) / self.warmup_t
            else:
                lr = self.lr_min + 0.5 * (base - self.lr_min) * (1 + math.cos(math.pi * epoch / self.t_initial))
            g["lr"] = lr


def create_scheduler(args, optimizer):
    epochs = args.epochs
    s = CosineLRScheduler(optimizer, t_initial=epochs, lr_min=getattr(args, "min_lr", 0.0),
                          warmup_t=getattr(args, "warmup_epochs", 0),
                          warmup_lr_init=getattr(args, "warmup_lr", 0.0))
    return s, epochs
