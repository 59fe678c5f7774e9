"""ORACLE (test infrastructure, NOT product code) -- restatement of loralib==0.1.2.

The reference depends on `loralib==0.1.2` (/root/reference/requirements.txt:2) which is
not vendored in the reference tree and is not installable here (no network).  This file
restates the published algorithm of microsoft/LoRA `loralib/layers.py` v0.1.2 for the
three symbols the hot path uses; call sites in the reference:
  - vit_pytorch_face/vit_face.py:330,333   lora.Linear(dim, hidden, r=lora_rank)
  - vit_pytorch_face/vit_face.py:349-355   lora.MergedLinear(..., r=0 [lora_pos "FFN"] or r=lora_rank [lora_pos "Attention"], enable_lora=[T,T,T], bias=False)
  - util/utils.py:573                      lora.Linear(in, out, r=rank)
  - train/train_own_forget_cl.py:316       lora.mark_only_lora_as_trainable(BACKBONE)

Parity status: PARITY UNPINNED at this boundary -- the reference holds no test, golden
vector or fixture for loralib; the restatement follows the published
semantics and is anchored on the reference's own call sites.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class LoRALayer:
    def __init__(self, r: int, lora_alpha: int, lora_dropout: float, merge_weights: bool):
        self.r = r
        self.lora_alpha = lora_alpha
        if lora_dropout > 0.0:
            self.lora_dropout = nn.Dropout(p=lora_dropout)
        else:
            self.lora_dropout = lambda x: x
        self.merged = False
        self.merge_weights = merge_weights


class Linear(nn.Linear, LoRALayer):
    """y = x W^T + b + (x A^T B^T) * (lora_alpha / r); W frozen when r > 0."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1,
                 lora_dropout: float = 0.0, fan_in_fan_out: bool = False,
                 merge_weights: bool = True, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                           merge_weights=merge_weights)
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
        self.reset_parameters()
        if fan_in_fan_out:
            self.weight.data = self.weight.data.transpose(0, 1)

    def reset_parameters(self):
        nn.Linear.reset_parameters(self)
        if hasattr(self, "lora_A"):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def _T(self, w):
        return w.transpose(0, 1) if self.fan_in_fan_out else w

    def train(self, mode: bool = True):
        nn.Linear.train(self, mode)
        if mode:
            if self.merge_weights and self.merged:
                if self.r > 0:
                    self.weight.data -= self._T(self.lora_B @ self.lora_A) * self.scaling
                self.merged = False
        else:
            if self.merge_weights and not self.merged:
                if self.r > 0:
                    self.weight.data += self._T(self.lora_B @ self.lora_A) * self.scaling
                self.merged = True
        return self

    def forward(self, x: torch.Tensor):
        if self.r > 0 and not self.merged:
            result = F.linear(x, self._T(self.weight), bias=self.bias)
            result = result + (self.lora_dropout(x) @ self.lora_A.transpose(0, 1)
                               @ self.lora_B.transpose(0, 1)) * self.scaling
            return result
        return F.linear(x, self._T(self.weight), bias=self.bias)


class MergedLinear(nn.Linear, LoRALayer):
    """One Linear whose output is `len(enable_lora)` equal slices (q | k | v), each enabled slice with its OWN rank-r pair:
    lora_A [r * n_enabled, in] stacks the A_g, lora_B [out / len * n_enabled, r] stacks the B_g, and the update of slice g is
    s * B_g A_g (upstream evaluates it as a grouped 1x1 conv1d of lora_A with lora_B, groups = n_enabled, scattered back into the
    enabled rows).  r = 0 is a plain frozen-or-not nn.Linear: what the reference builds with lora_pos == "FFN"
    (vit_face.py:349-355, 409-411); r > 0 is its lora_pos == "Attention" variant."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1,
                 lora_dropout: float = 0.0, enable_lora: List[bool] = [False],
                 fan_in_fan_out: bool = False, merge_weights: bool = True, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                           merge_weights=merge_weights)
        assert out_features % len(enable_lora) == 0, "The length of enable_lora must divide out_features"
        self.enable_lora = enable_lora
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0 and any(enable_lora):
            n_on = sum(enable_lora)
            self.lora_A = nn.Parameter(self.weight.new_zeros((r * n_on, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features // len(enable_lora) * n_on, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
            ind = self.weight.new_zeros((out_features,), dtype=torch.bool).view(len(enable_lora), -1)
            ind[enable_lora, :] = True
            self.lora_ind = ind.view(-1)
        self.reset_parameters()
        if fan_in_fan_out:
            self.weight.data = self.weight.data.transpose(0, 1)

    def reset_parameters(self):
        nn.Linear.reset_parameters(self)
        if hasattr(self, "lora_A"):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def _T(self, w):
        return w.transpose(0, 1) if self.fan_in_fan_out else w

    def merge_AB(self):
        """[out, in] update before scaling: rows of enabled slice g hold B_g A_g, disabled rows zero."""
        n_on = sum(self.enable_lora)
        delta = F.conv1d(self.lora_A.unsqueeze(0), self.lora_B.unsqueeze(-1), groups=n_on).squeeze(0)
        full = delta.new_zeros((len(self.lora_ind), *delta.shape[1:]))
        full[self.lora_ind] = delta
        return self._T(full)

    def train(self, mode: bool = True):
        nn.Linear.train(self, mode)
        live = self.r > 0 and any(self.enable_lora)
        if mode:
            if self.merge_weights and self.merged:
                if live:
                    self.weight.data -= self.merge_AB() * self.scaling
                self.merged = False
        else:
            if self.merge_weights and not self.merged:
                if live:
                    self.weight.data += self.merge_AB() * self.scaling
                self.merged = True
        return self

    def forward(self, x: torch.Tensor):
        result = F.linear(x, self._T(self.weight), bias=self.bias)
        if not self.merged and self.r > 0 and any(self.enable_lora):
            result = result + self.lora_dropout(x) @ self._T(self.merge_AB().T) * self.scaling
        return result


def mark_only_lora_as_trainable(model: nn.Module, bias: str = "none") -> None:
    for n, p in model.named_parameters():
        if "lora_" not in n:
            p.requires_grad = False
    if bias == "none":
        return
    if bias == "all":
        for n, p in model.named_parameters():
            if "bias" in n:
                p.requires_grad = True
    elif bias == "lora_only":
        for m in model.modules():
            if isinstance(m, LoRALayer) and hasattr(m, "bias") and m.bias is not None:
                m.bias.requires_grad = True
    else:
        raise NotImplementedError
