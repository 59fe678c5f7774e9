"""Engine plumbing shared by the engine-backed drop-in models (ViT_face, ModifiedViT): lazily creates the native engine,
keeps the LoRA nn.Parameters as views of the engine's flat fp32 buffer, refreshes the fp16 operand caches when a frozen weight
changed (load_state_dict, loralib merge / un-merge), hands out activation slots and wires the selective backward into autograd."""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import _ffi as F
from .engine import EngineSpec, VitEngine


class _EngineFn(torch.autograd.Function):
    """One autograd node for the whole network: inputs are the LoRA matrices, outputs (logits, emb)."""

    @staticmethod
    def forward(ctx, model, img, label, always_logits, *lora_params):
        eng = model._engine
        slot = model._take_slot()
        B = eng.forward(img, label, slot, use_lora=True, dropout_seed=model.dropout_seed(), **model.image_kwargs(img))
        ctx.model, ctx.slot, ctx.B, ctx.stamp = model, slot, B, model._slot_stamp[slot]
        emb = eng.slot_tensor(slot, F.SLOT_EMB, B).clone()
        if label is None and not always_logits:
            return emb
        return eng.slot_tensor(slot, F.SLOT_LOGITS, B).clone(), emb

    @staticmethod
    def backward(ctx, *grads):
        model, eng, slot = ctx.model, ctx.model._engine, ctx.slot
        if model._slot_stamp[slot] != ctx.stamp:
            raise RuntimeError("gslora-b200: the activations of this forward were overwritten by later forwards; raise "
                               "GSLORA_SLOTS (activation sets kept alive) or call backward sooner")
        if len(grads) == 2:
            dlogits, demb = grads
        else:
            dlogits, demb = None, grads[0]
        dlogits = dlogits.contiguous().float() if dlogits is not None else None
        demb = demb.contiguous().float() if demb is not None else None
        eng.backward(slot, dlogits, demb, accumulate=False)
        out = [eng.lora_view(eng.grad_flat, l, w).clone() for l in range(eng.spec.depth) for w in range(eng.spec.tensors_per_block)]
        return (None, None, None, None, *out)



class EngineBackedModel(nn.Module):
    """Subclasses provide: lora_layers() -> iterable of (fc1, fc2) loralib.Linear pairs per block, _frozen_tensors() -> the
    pointer table of include/gslora.h, engine_spec() -> EngineSpec, and the attributes dropout_p / emb_dropout_p."""

    def _init_engine_state(self):
        # engine state (not parameters / buffers: never part of state_dict)
        self._engine: Optional[VitEngine] = None
        self._frozen_sig = None
        self._lora_sig = None
        self._slot_next = 0
        self._slot_stamp: List[int] = []

    def __deepcopy__(self, memo):
        import copy
        # device-side companions are never copied: the engine (workspace), a captured CUDA graph of the step (static buffers, graph handle) and
        # the cached prototype table belong to THIS instance; the copy (EMA model, modify_head / resume_head clones) builds its own lazily
        skip = ("_engine", "_gsl_graph", "_gsl_proto_cache")
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in skip:
                setattr(new, k, copy.deepcopy(v, memo))
        new._engine, new._frozen_sig, new._lora_sig = None, None, None
        return new


    # uint8 input pipeline: raw pixels cross PCIe (1 byte / pixel instead of 4) and transforms.ToTensor() [+ Normalize] runs
    # inside the patchify kernel.  `input_pixel_norm = (mean, std)` mirrors transforms.Normalize of the ImageNet runs
    # (train_own_forget_cl.py:138-139); None = ToTensor only (the CASIA runs).
    input_pixel_norm = None

    def prepare_images(self, img: torch.Tensor) -> torch.Tensor:
        """fp32 NCHW (anything non-uint8 is cast, as the reference's modules would) or uint8 NCHW / NHWC, contiguous."""
        return img.contiguous() if img.dtype == torch.uint8 else img.float().contiguous()

    def image_kwargs(self, img: torch.Tensor) -> dict:
        if img.dtype != torch.uint8:
            return {}
        C = self.engine_spec().channels
        nhwc = img.dim() == 4 and img.shape[-1] == C and img.shape[1] != C
        return dict(pixel_norm=self.input_pixel_norm, channels_last=nhwc)

    def lora_parameters(self) -> List[nn.Parameter]:
        """flat-buffer order: per block, per LoRA layer of lora_layers() (fc1, fc2 -- or to_qkv alone with lora_pos "Attention"): lora_A, lora_B"""
        out = []
        for layers in self.lora_layers():
            for m in layers:
                out += [m.lora_A, m.lora_B]
        return out



    # Arithmetic of the frozen-weight GEMMs: None = GSLORA_PRECISION (default "split"), or "split" / "fast" per model
    # (include/gslora.h GslConfig.precision).  Changing it re-creates the engine at the next forward.
    gsl_precision: Optional[str] = None

    def _spec(self) -> EngineSpec:
        spec = self.engine_spec()
        if self.gsl_precision is not None:
            spec.precision = F.PRECISION_BY_NAME[self.gsl_precision]
        return spec

    def ensure_engine(self, batch: int, slots: Optional[int] = None) -> VitEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise F.GslError("gslora-b200: ViT_face executes on a CUDA device (sm_100a) only; there is no CPU fallback")
        slots = slots or int(os.environ.get("GSLORA_SLOTS", "2"))
        e = self._engine
        spec = self._spec()
        want_prec = F.default_precision() if spec.precision < 0 else spec.precision
        if e is None or e.device != dev or e.max_batch < batch or e.num_slots < slots or e.precision != want_prec:
            old = e
            same_dev = old is not None and old.device == dev
            cap = max(batch, old.max_batch if same_dev else 0)
            # the workspace grows, the training state must not change: LoRA values are re-linked by sync_engine, the fused AdamW moments and the
            # bias-correction step are carried over here (a fresh engine would silently restart Adam in the middle of a task)
            carry = None
            if old is not None and old.lora_flat.numel() == spec.depth * spec.lora_block_elems:
                carry = (old.exp_avg.detach().clone(), old.exp_avg_sq.detach().clone(), old.opt_step)
                old.workspace = None            # release the old activation workspace before the bigger one is allocated
            self._engine = None
            del old, e
            self._engine = VitEngine(spec, dev, cap, slots)
            if carry is not None:
                self._engine.exp_avg.copy_(carry[0].to(dev))
                self._engine.exp_avg_sq.copy_(carry[1].to(dev))
                self._engine.opt_step = carry[2]
            self._frozen_sig = self._lora_sig = None
            self._slot_stamp = [0] * slots
            self._slot_next = 0
        return self._engine

    def inference_chunk(self, batch: int) -> int:
        """Images per engine call of a no-grad forward.  The workspace is sized for TRAINING (every activation the selective backward needs, for
        max_batch x slots); an eval pass with a bigger loader batch (the drivers evaluate with 5x / 1000-image batches, train_own_forget_cl.py:
        899-937) must not grow it -- that would pin tens of GB for a forward that saves nothing -- so it runs in chunks of the current capacity
        (at least GSLORA_EVAL_CHUNK images, default 128, when no engine exists yet)."""
        floor = int(os.environ.get("GSLORA_EVAL_CHUNK", "128"))
        have = self._engine.max_batch if self._engine is not None else 0
        return max(1, min(batch, max(have, floor)))

    def sync_engine(self, force_lora: bool = False):
        """Re-link parameters into the engine and refresh its fp16 operand caches if anything changed."""
        eng = self._engine
        relinked = False
        for l, layers in enumerate(self.lora_layers()):
            for w, p in enumerate([q for m in layers for q in (m.lora_A, m.lora_B)]):
                view = eng.lora_view(eng.lora_flat, l, w)
                if p.data_ptr() != view.data_ptr():
                    view.copy_(p.data)
                    p.data = view
                    relinked = True
        params = self._frozen_tensors()
        frozen = [None if t is None else t.data for t in params]
        for t in frozen:
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise F.GslError("gslora-b200: frozen parameters must be contiguous fp32")
        sig = tuple((0, 0) if q is None else (q.data_ptr(), q._version) for q in params)
        sig += tuple(m._gsl_generation for pair in self.lora_layers() for m in pair)
        if sig != self._frozen_sig:
            eng.bind(frozen)
            eng.refresh_frozen()
            self._frozen_sig = sig
            self._lora_sig = None
        lsig = tuple(p._version for p in self.lora_parameters()) + (eng.opt_step,)
        if relinked or force_lora or lsig != self._lora_sig:
            eng.refresh_lora()
            self._lora_sig = lsig

    def mark_lora_updated_by_engine(self):
        self._lora_sig = tuple(p._version for p in self.lora_parameters()) + (self._engine.opt_step,)

    def _take_slot(self) -> int:
        s = self._slot_next
        self._slot_next = (s + 1) % self._engine.num_slots
        self._slot_stamp[s] += 1
        return s

    def dropout_seed(self) -> int:
        """Non-zero seed for the engine's counter-based dropout masks in train mode (drawn from torch's CPU generator, so
        torch.manual_seed makes runs reproducible); 0 (= no dropout) in eval mode or when both probabilities are 0."""
        if not self.training or (self.dropout_p <= 0.0 and self.emb_dropout_p <= 0.0) or os.environ.get("GSLORA_DROPOUT", "on") == "off":
            return 0
        return int(torch.randint(1, 2 ** 62, (1,)).item())

    def inference_slots(self, images: torch.Tensor, labels: Optional[torch.Tensor], dropout_seed: int = 0):
        """No-grad forward of a (possibly large) batch in chunks of the engine's capacity: yields (slot, lo, B) per chunk, results in the slot
        (eval_data counts hits, calculate_prototypes accumulates class sums).  `images` / `labels` already prepared (prepare_images, int64)."""
        chunk = self.inference_chunk(int(images.shape[0]))
        eng = self.ensure_engine(chunk)
        self.sync_engine()
        use_lora = not self._merged()
        for lo in range(0, int(images.shape[0]), chunk):
            slot = self._take_slot()
            lab = labels[lo:lo + chunk] if labels is not None else None
            B = eng.forward(images[lo:lo + chunk], lab, slot, use_lora=use_lora, dropout_seed=dropout_seed, **self.image_kwargs(images))
            yield slot, lo, B

    def _merged(self) -> bool:
        states = {m.merged for pair in self.lora_layers() for m in pair}
        if len(states) != 1:
            raise RuntimeError("gslora-b200: LoRA layers are in mixed merged / un-merged states")
        return states.pop()

    # ------------------------------------------------------------------ forward
    def _engine_forward(self, img, label=None, always_logits: bool = False):
        """(logits, emb) if `label` is given (or always_logits) else emb."""
        img = self.prepare_images(img)
        if label is not None:
            label = label.to(device=img.device, dtype=torch.int64).contiguous()
        merged = self._merged()
        lora_params = self.lora_parameters()
        need_grad = torch.is_grad_enabled() and not merged and any(p.requires_grad for p in lora_params)
        if need_grad:
            self.ensure_engine(img.shape[0])
            self.sync_engine()
            return _EngineFn.apply(self, img, label, always_logits, *lora_params)
        # no-grad forward: chunked to the engine's current capacity (see inference_chunk)
        embs, logits = [], []
        want_logits = label is not None or always_logits
        for slot, lo, B in self.inference_slots(img, label, dropout_seed=self.dropout_seed()):
            eng = self._engine
            embs.append(eng.slot_tensor(slot, F.SLOT_EMB, B).clone())
            if want_logits:
                logits.append(eng.slot_tensor(slot, F.SLOT_LOGITS, B).clone())
        emb = embs[0] if len(embs) == 1 else torch.cat(embs, dim=0)
        if not want_logits:
            return emb
        return (logits[0] if len(logits) == 1 else torch.cat(logits, dim=0)), emb
