"""Python handle on the native engine (csrc/gsl_engine.cu) -- plumbing only: owns the torch tensors the C library
borrows (workspace, flat LoRA parameter / gradient / optimizer-state buffers) and turns raw slot pointers back into
zero-copy torch views.  All arithmetic happens in libgslora.so; there is no PyTorch fallback."""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _ffi as F


@dataclass
class EngineSpec:
    image_size: int
    patch_size: int
    channels: int
    dim: int
    depth: int
    heads: int
    mlp_dim: int
    num_class: int
    lora_rank: int
    attn_scale: float
    ln_eps: float = 1e-5
    cos_s: float = 64.0
    cos_m: float = 0.35
    patch_order: int = 0
    grad_scale: float = 1024.0
    dropout: float = 0.0
    emb_dropout: float = 0.0
    head_type: int = 0          # 0 CosFace (ViT_face), 1 Linear + bias (torchvision heads.head)
    lora_pos: int = 0           # 0 "FFN" (lora.Linear on both FFN Linears), 1 "Attention" (lora.MergedLinear on to_qkv)
    precision: int = -1         # -1: GSLORA_PRECISION (default "split"); 0 fast (fp16 weights); 1 split (fp16 hi + lo weights, grads <= 1e-3)

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1

    @property
    def lora_block_elems(self) -> int:
        return sum(a * b for a, b in self.lora_shapes())

    def lora_shapes(self):
        r, D, H = self.lora_rank, self.dim, self.mlp_dim
        if self.lora_pos == 1:
            return [(3 * r, D), (3 * self.heads * 64, r)]     # to_qkv.lora_A (A_q | A_k | A_v), to_qkv.lora_B (B_q | B_k | B_v)
        return [(r, D), (H, r), (r, H), (D, r)]      # A(net.0) B(net.0) A(net.3) B(net.3)

    @property
    def tensors_per_block(self) -> int:
        return len(self.lora_shapes())


class VitEngine:
    _next_serial = 0

    def __init__(self, spec: EngineSpec, device: torch.device, max_batch: int, num_slots: int = 1):
        VitEngine._next_serial += 1
        self.serial = VitEngine._next_serial        # identity of this engine instance (id() can be recycled after a re-creation)
        if device.type != "cuda":
            raise F.GslError("gslora-b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.spec, self.device, self.max_batch, self.num_slots = spec, device, int(max_batch), int(num_slots)
        self.precision = F.default_precision() if spec.precision < 0 else int(spec.precision)
        self.cfg = F.GslConfig(image_size=spec.image_size, patch_size=spec.patch_size, channels=spec.channels, dim=spec.dim,
                               depth=spec.depth, heads=spec.heads, mlp_dim=spec.mlp_dim, num_class=spec.num_class,
                               lora_rank=spec.lora_rank, max_batch=self.max_batch, num_slots=self.num_slots,
                               patch_order=spec.patch_order, attn_scale=spec.attn_scale, ln_eps=spec.ln_eps, cos_s=spec.cos_s,
                               cos_m=spec.cos_m, lora_scaling=1.0 / spec.lora_rank, grad_scale=spec.grad_scale,
                               dropout=spec.dropout, emb_dropout=spec.emb_dropout, head_type=spec.head_type, precision=self.precision,
                               lora_pos=spec.lora_pos)
        L = F.lib()
        nbytes = L.gsl_engine_workspace_bytes(ctypes.byref(self.cfg))
        if nbytes == 0:
            raise F.GslError("unsupported configuration: " + L.gsl_last_error().decode())
        with torch.cuda.device(device):
            self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            pad = (-self.workspace.data_ptr()) % 1024
            self._ws_base = self.workspace.data_ptr() + pad
            self._ws_pad = pad
            h = ctypes.c_void_p()
            F.check(L.gsl_engine_create(ctypes.byref(self.cfg), ctypes.c_void_p(self._ws_base), nbytes, ctypes.byref(h)), "gsl_engine_create")
        self.handle = h
        n = spec.depth * spec.lora_block_elems
        self.lora_flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.grad_flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=device)
        self.opt_step = 0
        self.group_offsets = torch.arange(0, spec.depth + 1, dtype=torch.int32, device=device) * spec.lora_block_elems
        offs = []
        for l in range(spec.depth):
            o = l * spec.lora_block_elems
            for shp in spec.lora_shapes():
                offs.append(o)
                o += shp[0] * shp[1]
        offs.append(n)
        self.tensor_offsets_host = offs
        self.tensor_offsets = torch.tensor(offs, dtype=torch.int32, device=device)
        # group-lasso groupings of engine.get_structure_loss (engine.py:532-687): "block" = the 4 LoRA matrices of a block (what
        # engine_cl.get_structure_loss hard-codes, engine_cl.py:387-402), "lora" = one (A, B) pair, "matrix" = every matrix on its own.
        # All three are contiguous slices of the flat [A1 | B1 | A2 | B2] layout, so a grouping is just another offsets array.
        self.group_offsets_by_type = {
            "block": self.group_offsets,
            "lora": torch.tensor(offs[0::2], dtype=torch.int32, device=device),
            "matrix": self.tensor_offsets,
        }
        if spec.lora_pos == 1:      # engine.py:650-656: with LoRA on attention a group is the block's (lora_A, lora_B) pair whatever group_type says
            self.group_offsets_by_type = {k: self.group_offsets for k in ("block", "lora", "matrix")}
        self.group_norms = torch.zeros(4 * spec.depth, dtype=torch.float32, device=device)
        self.num_groups = spec.depth
        self.sums = torch.zeros(8, dtype=torch.float32, device=device)
        self._frozen_keep: List[torch.Tensor] = []

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                F.lib().gsl_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def lora_view(self, buf: torch.Tensor, block: int, which: int) -> torch.Tensor:
        o = self.tensor_offsets_host[self.spec.tensors_per_block * block + which]
        shp = self.spec.lora_shapes()[which]
        return buf[o:o + shp[0] * shp[1]].view(*shp)

    def bind(self, frozen: Sequence[Optional[torch.Tensor]]):
        exp = F.NUM_GLOBAL_PARAMS + F.NUM_BLOCK_PARAMS * self.spec.depth
        assert len(frozen) == exp, f"expected {exp} frozen tensors"
        arr = (ctypes.c_void_p * exp)()
        keep = []
        for i, t in enumerate(frozen):
            if t is None:
                arr[i] = None
                continue
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "frozen parameters must be contiguous fp32 CUDA tensors"
            arr[i] = t.data_ptr()
            keep.append(t)
        self._frozen_keep = keep
        F.check(F.lib().gsl_engine_bind_params(self.handle, arr, exp, F.ptr(self.lora_flat), F.ptr(self.grad_flat)), "gsl_engine_bind_params")

    def refresh_frozen(self):
        F.check(F.lib().gsl_engine_refresh_frozen(self.handle, F.cur_stream()), "gsl_engine_refresh_frozen")

    def refresh_lora(self):
        F.check(F.lib().gsl_engine_refresh_lora(self.handle, F.cur_stream()), "gsl_engine_refresh_lora")

    # ------------------------------------------------------------------ forward / backward
    def _view(self, ptr: int, shape, dtype) -> torch.Tensor:
        off = ptr - self.workspace.data_ptr()
        n = int(math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off:off + n].view(dtype).view(*shape)

    def slot_tensor(self, slot: int, what: int, B: int) -> torch.Tensor:
        p = F.lib().gsl_engine_slot_ptr(self.handle, slot, what)
        s = self.spec
        if what == F.SLOT_EMB:
            return self._view(p, (B, s.dim), torch.float32)
        if what == F.SLOT_LOGITS:
            return self._view(p, (B, s.num_class), torch.float32)
        if what == F.SLOT_CE:
            return self._view(p, (B,), torch.float32)
        if what == F.SLOT_CORRECT:
            return self._view(p, (B,), torch.int32)
        if what == F.SLOT_XFINAL:
            return self._view(p, (B, s.dim), torch.float32)
        raise ValueError(what)

    def forward(self, img: torch.Tensor, labels: Optional[torch.Tensor], slot: int = 0, use_lora: bool = True, dropout_seed: int = 0,
                pixel_norm=None, channels_last: bool = False, seed_dev: Optional[torch.Tensor] = None):
        """fp32 NCHW images (the reference loader's ToTensor output), or raw uint8 pixels ([B,C,S,S], or [B,S,S,C] with channels_last):
        ToTensor's /255 and the optional Normalize `pixel_norm = (mean, std)` are then applied inside the patchify kernel."""
        assert img.is_cuda and img.is_contiguous() and img.dtype in (torch.float32, torch.uint8)
        B = img.shape[0]
        if labels is not None:
            assert labels.is_cuda and labels.dtype == torch.int64 and labels.is_contiguous()
        if seed_dev is not None:
            # CUDA-graph form: the dropout base seed is read on the device from the step-state block (gsl_engine_forward_dev)
            assert pixel_norm is None, "the graph-capturable forward takes ToTensor output or raw pixels without Normalize"
            kind = 0 if img.dtype != torch.uint8 else (2 if channels_last else 1)
            F.check(F.lib().gsl_engine_forward_dev(self.handle, slot, F.ptr(img), kind, F.ptr(labels), B, 1 if use_lora else 0,
                                                   1 if dropout_seed else 0, F.ptr(seed_dev), F.cur_stream()), "gsl_engine_forward_dev")
            return B
        if img.dtype == torch.uint8:
            s = self.spec
            want = (B, s.image_size, s.image_size, s.channels) if channels_last else (B, s.channels, s.image_size, s.image_size)
            assert tuple(img.shape) == want, f"uint8 images {tuple(img.shape)} != {want}"
            mean, std = (F.host_floats(pixel_norm[0]), F.host_floats(pixel_norm[1])) if pixel_norm is not None else (None, None)
            F.check(F.lib().gsl_engine_forward_u8(self.handle, slot, F.ptr(img), 1 if channels_last else 0, mean, std, F.ptr(labels), B,
                                                  1 if use_lora else 0, int(dropout_seed), F.cur_stream()), "gsl_engine_forward_u8")
            return B
        assert pixel_norm is None and not channels_last, "pixel_norm / channels_last apply to uint8 images only"
        F.check(F.lib().gsl_engine_forward(self.handle, slot, F.ptr(img), F.ptr(labels), B, 1 if use_lora else 0, int(dropout_seed), F.cur_stream()),
                "gsl_engine_forward")
        return B

    def class_sums(self, slot: int, labels: torch.Tensor, B: int, sums: torch.Tensor, counts: torch.Tensor):
        """calculate_prototypes' accumulation (util/utils.py:535-541) for the slot's embeddings: sums[label] += emb, counts[label] += 1."""
        emb = self.slot_tensor(slot, F.SLOT_EMB, B)
        F.check(F.lib().gsl_class_sums(F.ptr(emb), F.ptr(labels), B, self.spec.dim, sums.shape[0], F.ptr(sums), F.ptr(counts), F.cur_stream()),
                "gsl_class_sums")

    def backward(self, slot: int, dlogits: Optional[torch.Tensor], demb: Optional[torch.Tensor], accumulate: bool = False):
        for t in (dlogits, demb):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        F.check(F.lib().gsl_engine_backward(self.handle, slot, F.ptr(dlogits), F.ptr(demb), 1 if accumulate else 0, F.cur_stream()),
                "gsl_engine_backward")

    # ------------------------------------------------------------------ losses / optimizer (device side, no host sync)
    def loss_sums(self, slot: int, n_remain: int, B: int, kl: Optional[torch.Tensor] = None) -> torch.Tensor:
        ce = self.slot_tensor(slot, F.SLOT_CE, B)
        correct = self.slot_tensor(slot, F.SLOT_CORRECT, B)
        F.check(F.lib().gsl_loss_sums(F.ptr(ce), F.ptr(correct), F.ptr(kl), n_remain, B, F.ptr(self.sums), F.cur_stream()), "gsl_loss_sums")
        return self.sums

    def prototype_kl(self, slot: int, labels: torch.Tensor, proto: torch.Tensor, B: int) -> torch.Tensor:
        """Per-sample KL(log_softmax(emb) || log_softmax(proto[label])) of engine_cl.get_prototype_loss (engine_cl.py:571-603)."""
        emb = self.slot_tensor(slot, F.SLOT_EMB, B)
        kl = torch.empty(B, dtype=torch.float32, device=self.device)
        F.check(F.lib().gsl_prototype_kl_fwd(F.ptr(emb), F.ptr(labels), F.ptr(proto), B, self.spec.dim, F.ptr(kl), F.cur_stream()), "gsl_prototype_kl_fwd")
        return kl

    def prototype_kl_grad(self, slot: int, labels: torch.Tensor, proto: torch.Tensor, n_remain: int, B: int, w_f: float, w_r: float,
                          BND_pro: float) -> torch.Tensor:
        emb = self.slot_tensor(slot, F.SLOT_EMB, B)
        demb = torch.empty(B, self.spec.dim, dtype=torch.float32, device=self.device)
        F.check(F.lib().gsl_prototype_kl_grad(F.ptr(emb), F.ptr(labels), F.ptr(proto), F.ptr(self.sums), n_remain, B, self.spec.dim, float(w_f),
                                              float(w_r), float(BND_pro), F.ptr(demb), F.cur_stream()), "gsl_prototype_kl_grad")
        return demb

    def unlearn_ce_grad(self, slot: int, labels: torch.Tensor, n_remain: int, B: int, beta: float, BND: float, out: torch.Tensor):
        logits = self.slot_tensor(slot, F.SLOT_LOGITS, B)
        F.check(F.lib().gsl_unlearn_ce_grad(F.ptr(logits), F.ptr(labels), F.ptr(self.sums), n_remain, B, self.spec.num_class,
                                            float(beta), float(BND), F.ptr(out), F.cur_stream()), "gsl_unlearn_ce_grad")

    def optimizer_step(self, lr: float, wd: float, alpha: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0,
                       group_type: str = "block", state_dev: Optional[torch.Tensor] = None):
        """Fused group-Lasso + AdamW on the flat LoRA buffer; repacks the fp16 LoRA operands afterwards.  state_dev: the step count and lr are
        read from the device step state at run time (CUDA-graph capture; the caller advances `opt_step` per replay)."""
        self.opt_step += 1
        n = self.lora_flat.numel()
        offs = self.group_offsets_by_type[group_type]
        self.num_groups = offs.numel() - 1
        if state_dev is not None:
            F.check(F.lib().gsl_grouplasso_adamw_step_dev(F.ptr(self.lora_flat), F.ptr(self.grad_flat), F.ptr(self.exp_avg), F.ptr(self.exp_avg_sq),
                                                          F.ptr(offs), self.num_groups, n, float(wd), float(betas[0]), float(betas[1]), float(eps),
                                                          float(alpha), float(grad_scale), F.ptr(state_dev), F.ptr(self.group_norms),
                                                          F.cur_stream()), "gsl_grouplasso_adamw_step_dev")
            self.refresh_lora()
            return
        F.check(F.lib().gsl_grouplasso_adamw_step(F.ptr(self.lora_flat), F.ptr(self.grad_flat), F.ptr(self.exp_avg), F.ptr(self.exp_avg_sq),
                                                  F.ptr(offs), self.num_groups, n, float(lr), float(wd), float(betas[0]),
                                                  float(betas[1]), float(eps), float(alpha), float(grad_scale), self.opt_step,
                                                  F.ptr(self.group_norms), F.cur_stream()), "gsl_grouplasso_adamw_step")
        self.refresh_lora()

    def reset_optimizer(self):
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.opt_step = 0

    def tensor_norms(self, type: str = "L2") -> torch.Tensor:
        n = self.spec.tensors_per_block * self.spec.depth
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        F.check(F.lib().gsl_tensor_norms(F.ptr(self.lora_flat), F.ptr(self.tensor_offsets), n, 0 if type == "L2" else 1,
                                         F.ptr(out), F.cur_stream()), "gsl_tensor_norms")
        return out
