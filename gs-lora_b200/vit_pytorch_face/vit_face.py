"""gslora-b200 `ViT_face` -- the reference's module surface (vit_pytorch_face/vit_face.py:449-548 of bjzhb666/GS-LoRA)
on top of the native sm_100a engine.

The module tree, constructor keywords, `forward(img, label=None, mask=None)` contract and every `state_dict` key are
the reference's, so `train/train_own_forget_cl.py` can build, load, freeze, wrap and checkpoint this
model unchanged.  The sub-modules are *parameter holders*: `ViT_face.forward` hands the whole computation
(patch-embed -> [LN, QKV, attention, out-proj, LN, fc1+LoRA, GELU, fc2+LoRA] x depth -> LN -> CosFace) to
libgslora.so, and the backward is the engine's selective backward wired in through one `torch.autograd.Function`
whose only differentiable leaves are the LoRA matrices.  There is no PyTorch fallback.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn

import loralib as lora
from gslora import _ffi as F
from gslora.engine import EngineSpec
from gslora.model_base import EngineBackedModel

MIN_NUM_PATCHES = 16


def _holder_forward(self, *a, **k):
    raise RuntimeError(f"gslora-b200: {type(self).__name__} is a parameter holder; it is executed by the fused engine via ViT_face.forward")


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn
    forward = _holder_forward


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn
    forward = _holder_forward


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0, lora_rank=8):
        super().__init__()
        self.net = nn.Sequential(lora.Linear(dim, hidden_dim, r=lora_rank), nn.GELU(), nn.Dropout(dropout),
                                 lora.Linear(hidden_dim, dim, r=lora_rank), nn.Dropout(dropout))
    forward = _holder_forward


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, lora_rank=0):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim ** -0.5          # the reference scales by dim, not dim_head (vit_face.py:346)
        self.to_qkv = lora.MergedLinear(in_features=dim, out_features=inner_dim * 3, r=lora_rank, enable_lora=[True, True, True], bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
    forward = _holder_forward


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout, lora_rank, up=False, lora_pos: str = "FFN"):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                Residual(PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout,
                                                lora_rank=(lora_rank if lora_pos == "Attention" else 0)))),
                Residual(PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout, lora_rank=lora_rank if lora_pos == "FFN" else 0))),
            ]))
        self.up = up
        self.depth = depth
    forward = _holder_forward


class CosFace(nn.Module):
    """s * (cos(theta) - m * onehot)  (vit_face.py:146-208).  Parameter holder: evaluated by the engine's head kernel."""

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.35):
        super().__init__()
        self.in_features, self.out_features, self.device_id, self.s, self.m = in_features, out_features, device_id, s, m
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
    forward = _holder_forward

    def __repr__(self):
        return f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, s={self.s}, m={self.m})"


class ViT_face(EngineBackedModel):
    def __init__(self, *, loss_type, GPU_ID, num_class, image_size, patch_size, dim, depth, heads, mlp_dim, pool="cls", channels=3,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, lora_rank=8, lora_pos: str = "FFN"):
        super().__init__()
        assert image_size % patch_size == 0, "Image dimensions must be divisible by the patch size."
        num_patches = (image_size // patch_size) ** 2
        patch_dim = channels * patch_size ** 2
        assert num_patches > MIN_NUM_PATCHES, f"your number of patches ({num_patches}) is way too small for attention to be effective (at least 16). Try decreasing your patch size"
        assert pool in {"cls", "mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        if pool != "cls" or lora_pos not in ("FFN", "Attention") or dim_head != 64 or heads * dim_head != dim:
            raise NotImplementedError("gslora-b200 builds the GS-LoRA configurations: pool='cls', lora_pos 'FFN' or 'Attention', dim_head=64, heads*64 == dim")
        self.lora_pos = lora_pos
        self.patch_size, self.image_size, self.channels = patch_size, image_size, channels
        self.dim, self.depth, self.heads, self.mlp_dim, self.num_class, self.lora_rank = dim, depth, heads, mlp_dim, num_class, lora_rank
        self.dropout_p, self.emb_dropout_p = float(dropout), float(emb_dropout)

        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.patch_to_embedding = nn.Linear(patch_dim, dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout, lora_rank, lora_pos=lora_pos)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim))
        self.loss_type = loss_type
        self.GPU_ID = GPU_ID
        if loss_type == "None":
            print("no loss for vit_face")
        elif loss_type == "CosFace":
            self.loss = CosFace(in_features=dim, out_features=num_class, device_id=GPU_ID)
        else:
            raise NotImplementedError(f"gslora-b200 builds the CosFace head only (every GS-LoRA script uses -head CosFace); got {loss_type}")
        self._init_engine_state()

    def lora_layers(self):
        """per block, the layers that carry LoRA: (fc1, fc2) for lora_pos "FFN", (to_qkv,) for lora_pos "Attention" (vit_face.py:405-425)"""
        for attn, ff in self.transformer.layers:
            if self.lora_pos == "Attention":
                yield (attn.fn.fn.to_qkv,)
            else:
                yield ff.fn.fn.net[0], ff.fn.fn.net[3]

    def _frozen_tensors(self):
        t = [self.pos_embedding, self.cls_token, self.patch_to_embedding.weight, self.patch_to_embedding.bias,
             self.mlp_head[0].weight, self.mlp_head[0].bias, self.loss.weight if hasattr(self, "loss") else None, None]
        for attn, ff in self.transformer.layers:
            a, f = attn.fn, ff.fn
            t += [a.norm.weight, a.norm.bias, a.fn.to_qkv.weight, a.fn.to_qkv.bias, a.fn.to_out[0].weight, a.fn.to_out[0].bias,
                  f.norm.weight, f.norm.bias, f.fn.net[0].weight, f.fn.net[0].bias, f.fn.net[3].weight, f.fn.net[3].bias]
        return t

    def engine_spec(self) -> EngineSpec:
        return EngineSpec(image_size=self.image_size, patch_size=self.patch_size, channels=self.channels, dim=self.dim, depth=self.depth,
                          heads=self.heads, mlp_dim=self.mlp_dim, num_class=self.num_class, lora_rank=self.lora_rank,
                          attn_scale=self.dim ** -0.5, ln_eps=self.mlp_head[0].eps,
                          cos_s=getattr(getattr(self, "loss", None), "s", 64.0), cos_m=getattr(getattr(self, "loss", None), "m", 0.35),
                          grad_scale=float(os.environ.get("GSLORA_GRAD_SCALE", "1024")), dropout=self.dropout_p, emb_dropout=self.emb_dropout_p,
                          lora_pos=1 if self.lora_pos == "Attention" else 0)

    def forward(self, img, label=None, mask=None):
        """:return: (logits, emb) if `label` is given else emb -- as vit_face.py:523-548"""
        if mask is not None:
            raise NotImplementedError("gslora-b200: attention masks are not used by any GS-LoRA script and are not built")
        if label is not None and not hasattr(self, "loss"):
            raise RuntimeError("labelled forward needs loss_type='CosFace'")
        return self._engine_forward(img, label)
