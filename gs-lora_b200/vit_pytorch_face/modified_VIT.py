"""gslora-b200 `ModifiedViT` -- the reference's torchvision wrapper (vit_pytorch_face/modified_VIT.py:5-39) on the native engine.

    model = ModifiedViT(torchvision.models.vit_b_16(weights=...))      # train/train_own_forget_cl.py:226-242
    util.utils.replace_ffn_with_lora(model, rank=8)                     # mlp.0 / mlp.3 -> loralib.Linear (util/utils.py:552-576)
    logits, cls_emb = model(x, label)                                   # label unused, as in the reference

Attribute names (conv_proj, class_token, encoder, heads) and therefore every state_dict key are the reference's.  The torchvision
sub-modules are parameter holders; the arithmetic (Conv-patchify as a GEMM with (c p1 p2) patch vectors, pos-embedding add,
pre-LN blocks with biased in_proj and 1/sqrt(64) attention scale, LN eps 1e-6, Linear head) runs in libgslora.so.
Requirements of this build: hidden_dim = heads * 64 (ViT-B/16, ViT-L/16 and the test-size variants), every encoder MLP Linear
replaced by loralib.Linear with the same rank (8 or 16), dropout 0 (torchvision default)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

import loralib as lora
from gslora.engine import EngineSpec
from gslora.model_base import EngineBackedModel


class ModifiedViT(EngineBackedModel):
    def __init__(self, vit_model):
        super().__init__()
        self.conv_proj = vit_model.conv_proj
        self.class_token = vit_model.class_token
        self.encoder = vit_model.encoder
        self.heads = vit_model.heads
        self.image_size, self.patch_size, self.hidden_dim = vit_model.image_size, vit_model.patch_size, vit_model.hidden_dim
        self.dropout_p, self.emb_dropout_p = float(getattr(vit_model, "dropout", 0.0)), float(getattr(vit_model, "dropout", 0.0))
        self._init_engine_state()

    def _blocks(self):
        return list(self.encoder.layers.children())

    def lora_layers(self):
        for blk in self._blocks():
            fc1, fc2 = blk.mlp[0], blk.mlp[3]
            if not (isinstance(fc1, lora.Linear) and isinstance(fc2, lora.Linear) and fc1.r > 0 and fc1.r == fc2.r):
                raise RuntimeError("gslora-b200 ModifiedViT: call util.utils.replace_ffn_with_lora(model, rank) first "
                                   "(every encoder MLP Linear must be a loralib.Linear of one rank)")
            yield fc1, fc2

    def _frozen_tensors(self):
        D = self.hidden_dim
        head = self.heads.head
        t = [self.encoder.pos_embedding, self.class_token, self.conv_proj.weight, self.conv_proj.bias, self.encoder.ln.weight, self.encoder.ln.bias,
             head.weight, head.bias]
        for blk in self._blocks():
            sa = blk.self_attention
            t += [blk.ln_1.weight, blk.ln_1.bias, sa.in_proj_weight, sa.in_proj_bias, sa.out_proj.weight, sa.out_proj.bias,
                  blk.ln_2.weight, blk.ln_2.bias, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[3].weight, blk.mlp[3].bias]
        return t

    def engine_spec(self) -> EngineSpec:
        blk = self._blocks()[0]
        heads = blk.self_attention.num_heads
        D = self.hidden_dim
        if D != heads * 64:
            raise NotImplementedError("gslora-b200: the engine's attention kernels are built for head_dim 64")
        return EngineSpec(image_size=self.image_size, patch_size=self.patch_size, channels=self.conv_proj.in_channels, dim=D,
                          depth=len(self._blocks()), heads=heads, mlp_dim=blk.mlp[0].out_features, num_class=self.heads.head.out_features,
                          lora_rank=blk.mlp[0].r, attn_scale=64 ** -0.5, ln_eps=self.encoder.ln.eps, patch_order=1, head_type=1,
                          grad_scale=float(os.environ.get("GSLORA_GRAD_SCALE", "1024")), dropout=self.dropout_p, emb_dropout=self.emb_dropout_p)

    def forward(self, x, label=None):
        # label is not used in this model (modified_VIT.py:24); passing it lets the head kernel also emit CE / hit counts
        return self._engine_forward(x, label, always_logits=True)
