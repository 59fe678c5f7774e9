import math
from typing import Dict, List

import torch
import torch.nn as nn


class LoRALayer:
    """Book-keeping mixin: rank, alpha, merge state.  `_gsl_generation` is bumped whenever the frozen weight is
    mutated (merge / un-merge) so the engine's fp16 operand caches know they are stale."""

    def __init__(self, r: int, lora_alpha: int, lora_dropout: float, merge_weights: bool):
        self.r = r
        self.lora_alpha = lora_alpha
        self.lora_dropout_p = float(lora_dropout)
        self.lora_dropout = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else (lambda x: x)
        self.merged = False
        self.merge_weights = merge_weights
        self._gsl_generation = 0


class Linear(nn.Linear, LoRALayer):
    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.0,
                 fan_in_fan_out: bool = False, merge_weights: bool = True, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r, lora_alpha, lora_dropout, merge_weights)
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

    def delta_weight(self) -> torch.Tensor:
        d = (self.lora_B.data @ self.lora_A.data) * self.scaling
        return d.transpose(0, 1) if self.fan_in_fan_out else d

    def train(self, mode: bool = True):
        nn.Linear.train(self, mode)
        if not self.merge_weights or self.r <= 0:
            return self
        if mode and self.merged:            # back to training: W -= B A * scaling
            self.weight.data -= self.delta_weight()
            self.merged = False
            self._gsl_generation += 1
        elif not mode and not self.merged:  # eval: W += B A * scaling
            self.weight.data += self.delta_weight()
            self.merged = True
            self._gsl_generation += 1
        return self

    def forward(self, x: torch.Tensor):
        from gslora.lora_ops import lora_linear_forward
        return lora_linear_forward(self, x)


class MergedLinear(nn.Linear, LoRALayer):
    """loralib.MergedLinear (0.1.2): one Linear whose output is len(enable_lora) equal slices (q | k | v), each enabled slice g with its own
    rank-r pair -- lora_A [r * n_on, in] stacks the A_g, lora_B [out / len * n_on, r] stacks the B_g, slice g's update is s * B_g A_g.
    The reference builds it with r = 0 for lora_pos "FFN" (a plain bias-free Linear) and with r = lora_rank, enable_lora = [True] * 3 for
    lora_pos "Attention" (vit_face.py:349-355, 405-425).  Parameter holder: the engine folds the update into its cached to_qkv operand."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.0,
                 enable_lora: List[bool] = [False], fan_in_fan_out: bool = False, merge_weights: bool = True, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r, lora_alpha, lora_dropout, merge_weights)
        assert out_features % len(enable_lora) == 0, "The length of enable_lora must divide out_features"
        self.enable_lora = list(enable_lora)
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0 and any(enable_lora):
            if not all(enable_lora) or fan_in_fan_out:
                raise NotImplementedError("gslora-b200: MergedLinear is built for enable_lora = [True] * n, fan_in_fan_out = False (the reference's to_qkv)")
            n_on = sum(enable_lora)
            self.lora_A = nn.Parameter(self.weight.new_zeros((r * n_on, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features // len(enable_lora) * n_on, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
        self.reset_parameters()

    def reset_parameters(self):
        nn.Linear.reset_parameters(self)
        if hasattr(self, "lora_A"):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def delta_weight(self) -> torch.Tensor:
        """[out, in]: rows of slice g = B_g A_g * scaling (upstream's merge_AB, a grouped 1x1 conv1d, times scaling)"""
        n, r = len(self.enable_lora), self.r
        per = self.out_features // n
        return torch.cat([self.lora_B.data[g * per:(g + 1) * per] @ self.lora_A.data[g * r:(g + 1) * r] for g in range(n)], dim=0) * self.scaling

    def train(self, mode: bool = True):
        nn.Linear.train(self, mode)
        if not self.merge_weights or self.r <= 0 or not any(self.enable_lora):
            return self
        if mode and self.merged:
            self.weight.data -= self.delta_weight()
            self.merged = False
            self._gsl_generation += 1
        elif not mode and not self.merged:
            self.weight.data += self.delta_weight()
            self.merged = True
            self._gsl_generation += 1
        return self

    def forward(self, x: torch.Tensor):
        from gslora.lora_ops import lora_linear_forward
        return lora_linear_forward(self, x)


def mark_only_lora_as_trainable(model: nn.Module, bias: str = "none") -> None:
    for name, p in model.named_parameters():
        if "lora_" not in name:
            p.requires_grad = False
    if bias == "none":
        return
    if bias == "all":
        for name, p in model.named_parameters():
            if "bias" in name:
                p.requires_grad = True
    elif bias == "lora_only":
        for m in model.modules():
            if isinstance(m, LoRALayer) and getattr(m, "bias", None) is not None:
                m.bias.requires_grad = True
    else:
        raise NotImplementedError(bias)


def lora_state_dict(model: nn.Module, bias: str = "none") -> Dict[str, torch.Tensor]:
    sd = model.state_dict()
    if bias == "none":
        return {k: v for k, v in sd.items() if "lora_" in k}
    if bias == "all":
        return {k: v for k, v in sd.items() if "lora_" in k or "bias" in k}
    raise NotImplementedError(bias)
