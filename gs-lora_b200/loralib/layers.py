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
    """The reference only builds this with r = 0 on the GS-LoRA path (`lora_pos == "FFN"`, vit_face.py:349-355,409-411):
    a plain frozen bias-free Linear.  r > 0 (LoRA on attention) is not built yet."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.0,
                 enable_lora: List[bool] = [False], fan_in_fan_out: bool = False, merge_weights: bool = True, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r, lora_alpha, lora_dropout, merge_weights)
        if r > 0:
            raise NotImplementedError("gslora-b200: MergedLinear with r > 0 (LoRA on attention, lora_pos='Attention') is not built yet")
        self.enable_lora = enable_lora
        self.fan_in_fan_out = fan_in_fan_out

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
